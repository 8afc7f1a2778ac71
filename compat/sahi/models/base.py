from fsd_b200.sahi_api.base import DetectionModel  # noqa: F401
