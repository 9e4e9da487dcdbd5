from .hypernet import HyperStructure  # noqa: F401
from .unet import UNet2DConditionModel, UNet2DConditionModelGated, UNet2DConditionModelPruned  # noqa: F401
from .encoders import AutoencoderKL, CLIPTextModel  # noqa: F401
