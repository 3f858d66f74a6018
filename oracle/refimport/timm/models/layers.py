from models.layers import *  # noqa: re-export of the reference's vendored copy
