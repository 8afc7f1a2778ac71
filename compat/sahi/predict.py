from fsd_b200.sahi_api.predict import (LOW_MODEL_CONFIDENCE, POSTPROCESS_NAME_TO_CLASS, filter_predictions,  # noqa: F401
                                       get_prediction, get_sliced_prediction)
