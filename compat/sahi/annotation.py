from fsd_b200.sahi_api.annotation import BoundingBox, Category, ObjectAnnotation  # noqa: F401
