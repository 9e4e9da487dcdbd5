from .trainer import (BilevelUnetFineTuner, ConstantWithWarmup, FusedAdamW, GradReducer, NoiseScheduler,  # noqa: F401
                      UnetFineTuner, cast_block_act_hooks, encode_prompt, fused_kd_loss)
