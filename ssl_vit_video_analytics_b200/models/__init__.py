from .tiny_vit import TinyViT, tiny_vit_21m_variant  # noqa: F401
from .mae_vit_adapter import TinyVideoMAE  # noqa: F401
from .videomae_vit import VideoMAEViT, VideoMAEViTClassifier  # noqa: F401
from .video_classifier import TinyViTVideoClassifier, apply_finetune_mode, resolve_mode  # noqa: F401
