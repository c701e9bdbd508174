"""Import side effects: register every module of the path in MODELS / POSITIONAL_ENCODING."""
from . import depther, hahi, heads, swin  # noqa: F401
from .builder import (BACKBONES, DEPTHER, HEADS, LOSSES, MODELS, NECKS, build_backbone,  # noqa: F401
                      build_depther, build_head, build_loss, build_neck)
