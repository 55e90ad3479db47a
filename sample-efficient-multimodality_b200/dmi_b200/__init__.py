"""dmi_b200 -- B200-native (sm_100a) implementation of the adapted-projector hot path of
ospanbatyr/sample-efficient-multimodality.  Python host code mirroring ``dmi/model`` over a C-ABI CUDA library."""
__version__ = "0.1.0"
