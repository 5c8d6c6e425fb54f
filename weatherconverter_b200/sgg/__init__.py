from .sgg import apply_gsg, apply_gsg_batch  # noqa: F401
