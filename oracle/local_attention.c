/* ORACLE -- TEST INFRASTRUCTURE ONLY (never linked or called by the product path).
 *
 * Plain-C restatement of the two forward operators of the external `localAttention`
 * extension (zzd1992/Image-Local-Attention @ master, un-vendored; requirements.txt:7),
 * as called from /root/reference/model/attention.py:18 (similar_forward) and :38
 * (weighting_forward).  Layout and tap order follow the reference's own
 * f_weighting_cpu (model/attention.py:75-85): taps row-major (i over kH, j over kW),
 * zero padding (an out-of-image tap contributes logit 0 / value 0).
 *
 * similar_forward has no working in-repo restatement -> "parity unpinned" (see
 * oracle/arseg_oracle.py header).
 */
#include <stddef.h>

/* S[n,y,x,i*kW+j] = sum_c Q[n,c,y,x] * K[n,c,y+i-kH/2,x+j-kW/2]
 * [lo,hi) is a range of flattened (n,y) rows: the Python caller splits it over threads. */
void oracle_similar_forward(const float* q, const float* k, float* out,
                            int N, int C, int H, int W, int kH, int kW, int lo, int hi) {
    const int rh = kH / 2, rw = kW / 2, T = kH * kW;
    const size_t plane = (size_t)H * W;
    for (int r = lo; r < hi && r < N * H; ++r) {
        const int n = r / H, y = r % H;
        for (int x = 0; x < W; ++x) {
            float* o = out + (((size_t)n * H + y) * W + x) * T;
            for (int i = 0; i < kH; ++i)
                for (int j = 0; j < kW; ++j) {
                    const int yy = y + i - rh, xx = x + j - rw;
                    float acc = 0.f;
                    if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
                        const float* qp = q + (size_t)n * C * plane + (size_t)y * W + x;
                        const float* kp = k + (size_t)n * C * plane + (size_t)yy * W + xx;
                        for (int c = 0; c < C; ++c) acc += qp[c * plane] * kp[c * plane];
                    }
                    o[i * kW + j] = acc;
                }
        }
    }
}

/* O[n,c,y,x] = sum_ij V[n,c,y+i-r,x+j-r] * A[n,y,x,i*kW+j]
 * [lo,hi) is a range of flattened (n,c) planes. */
void oracle_weighting_forward(const float* v, const float* a, float* out,
                              int N, int C, int H, int W, int kH, int kW, int lo, int hi) {
    const int rh = kH / 2, rw = kW / 2, T = kH * kW;
    const size_t plane = (size_t)H * W;
    for (int r = lo; r < hi && r < N * C; ++r) {
        const int n = r / C, c = r % C;
        const float* vp = v + ((size_t)n * C + c) * plane;
        float* op = out + ((size_t)n * C + c) * plane;
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                const float* ap = a + (((size_t)n * H + y) * W + x) * T;
                float acc = 0.f;
                for (int i = 0; i < kH; ++i) {
                    const int yy = y + i - rh;
                    if (yy < 0 || yy >= H) continue;
                    for (int j = 0; j < kW; ++j) {
                        const int xx = x + j - rw;
                        if (xx < 0 || xx >= W) continue;
                        acc += vp[(size_t)yy * W + xx] * ap[i * kW + j];
                    }
                }
                op[(size_t)y * W + x] = acc;
            }
    }
}
