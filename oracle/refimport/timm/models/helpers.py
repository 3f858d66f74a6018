from models.helpers import *  # noqa: re-export of the reference's vendored copy
from models.helpers import build_model_with_cfg, overlay_external_default_cfg  # noqa
