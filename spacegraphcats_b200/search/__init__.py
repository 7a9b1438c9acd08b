from .catlas import CAtlas  # noqa: F401
