from .annotation import RectangleAnnotation  # noqa: F401
from .image import ImageMetadata  # noqa: F401
from .project import Label, Project  # noqa: F401
