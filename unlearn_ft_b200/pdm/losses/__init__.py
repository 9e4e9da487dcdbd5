from .contrastive_loss import ContrastiveLoss  # noqa: F401
from .resource_loss import ResourceLoss  # noqa: F401
