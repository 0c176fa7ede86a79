from .ip_adapter import IPAdapter, IPAdapterPlus, ImageProjModel  # noqa: F401
from .attention_processor import Attention, AttnProcessor, IPAttnProcessor  # noqa: F401
from .resampler import Resampler  # noqa: F401
