from .unet_2d_conditional import (UNet2DConditionModel, UNet2DConditionModelGated, UNet2DConditionModelPruned,  # noqa
                                  UNet2DConditionOutput)
