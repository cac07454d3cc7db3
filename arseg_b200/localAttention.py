"""`localAttention` operator API (the 5 functions imported at /root/reference/model/attention.py:7-11)."""
from .ops import (similar_forward, similar_backward, weighting_forward,  # noqa: F401
                  weighting_backward_ori, weighting_backward_weight)
