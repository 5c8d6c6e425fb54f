from .models import Generator  # noqa: F401
