"""Pin the oracle's merge_motion against the UNMODIFIED reference function
(pre-process/generate_compressed_dataset_camvid.py:6-56), run in the authoring container:

    python tests/golden/make_merge_motion_golden.py

The reference reads `000.png` (for the frame size) and `test_%03d.bin` (hard-coded 720x960x3 int16) from a workspace
directory, so the synthetic decoder maps (arseg_b200.synth.synth_decoder_maps, seeds recorded) are written there first.
Writes tests/golden/merge_motion.npz: the merged MV field of the last frame (int16 [720,960,2]) and a CRC of every plane.
"""
import importlib.util
import os
import sys
import tempfile
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from arseg_b200 import synth  # noqa: E402
from oracle import arseg_oracle as O  # noqa: E402

REF = "/root/reference/pre-process/generate_compressed_dataset_camvid.py"
F, H, W, SEED = 4, 720, 960, 11


def main():
    import cv2
    src = open(REF).read()
    # only the function under test: the module body below it walks the dataset on disk
    head = src[:src.index("scene_length_info")]
    mod = {}
    exec(compile(head, REF, "exec"), mod)
    maps = synth.synth_decoder_maps(F, H, W, SEED)
    with tempfile.TemporaryDirectory() as d:
        cv2.imwrite(os.path.join(d, "000.png"), np.zeros((H, W, 3), np.uint8))
        for f in range(1, F + 1):
            maps[f - 1].tofile(os.path.join(d, "test_%03d.bin" % f))
        ref = mod["mergeMotion"](d, 0, F)
    mine = O.merge_motion(maps)
    assert ref.shape == mine.shape == (H, W, F + 1, 2), (ref.shape, mine.shape)
    assert np.array_equal(ref, mine), "oracle.merge_motion differs from the reference"
    crcs = [zlib.crc32(np.ascontiguousarray(ref[:, :, f].astype(np.int16)).tobytes()) for f in range(F + 1)]
    np.savez_compressed(os.path.join(HERE, "merge_motion.npz"), last=ref[:, :, F].astype(np.int16), crcs=np.array(crcs, dtype=np.int64),
                        F=F, H=H, W=W, seed=SEED)
    print("oracle == reference on %d frames of %dx%d; golden written (%d bytes)" %
          (F, H, W, os.path.getsize(os.path.join(HERE, "merge_motion.npz"))))


if __name__ == "__main__":
    main()
