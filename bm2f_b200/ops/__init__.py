"""Mirror of the reference's `mask2former/modeling/pixel_decoder/ops` package
(functions/ + modules/), backed by the sm_100a extension."""
from .functions import MSDeformAttnFunction  # noqa: F401
from .modules import MSDeformAttn  # noqa: F401
