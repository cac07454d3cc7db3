"""Import-compatible replacement of the `localAttention` pip extension
(git+https://github.com/zzd1992/Image-Local-Attention.git, requirements.txt:7) that the reference imports at
model/attention.py:7-11.  Put this directory's parent on PYTHONPATH *instead of* installing the upstream
package; the unmodified reference then runs its CReFF ops on the sm_100a kernels of libarseg_sm100a.so.
"""
from arseg_b200.ops import (similar_forward, similar_backward, weighting_forward,  # noqa: F401
                            weighting_backward_ori, weighting_backward_weight)

__all__ = ["similar_forward", "similar_backward", "weighting_forward", "weighting_backward_ori",
           "weighting_backward_weight"]
