from .sgg import apply_gsg, apply_gsg_batch, apply_lcg  # noqa: F401
