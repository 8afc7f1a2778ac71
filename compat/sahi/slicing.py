from fsd_b200.sahi_api.slicing import SliceImageResult, get_slice_bboxes, slice_image  # noqa: F401
