from models.vision_transformer import *  # noqa: re-export of the reference's vendored copy
from models.vision_transformer import _cfg, Mlp, Attention, Block, PatchEmbed  # noqa
