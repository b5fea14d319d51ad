"""Import-compatibility shim: the reference keeps its counter in `model/utils/parm_octconv_v2.py`."""
from .simplesum_octconv import print_model_parm_flops, print_model_parm_nums  # noqa: F401
