"""Drop-in mirror of the SAHI surface the reference uses: same names, arguments and error behaviour as
docs sahi/{predict,prediction,base}.py and the external sahi.annotation / sahi.slicing / sahi.postprocess.combine,
with the arithmetic routed to the sm_100a kernels.  `compat/sahi` re-exports these under the upstream module paths."""
from .annotation import BoundingBox, Category, ObjectAnnotation  # noqa: F401
from .prediction import ObjectPrediction, PredictionResult, PredictionScore  # noqa: F401
from .slicing import get_slice_bboxes, read_image_as_pil, slice_image  # noqa: F401
from .base import DetectionModel  # noqa: F401
from .postprocess import (GreedyNMMPostprocess, LSNMSPostprocess, NMMPostprocess, NMSPostprocess,  # noqa: F401
                          PostprocessPredictions)
from .predict import get_prediction, get_sliced_prediction  # noqa: F401
