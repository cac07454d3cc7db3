"""arseg_b200 -- B200-native (sm_100a) implementation of AR-Seg's per-non-keyframe inference path."""
__version__ = "0.1.0"
