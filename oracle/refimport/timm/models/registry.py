from models.registry import *  # noqa: re-export of the reference's vendored copy
from models.registry import register_model  # noqa
