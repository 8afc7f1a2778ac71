from fsd_b200.sahi_api.prediction import ObjectPrediction, PredictionResult, PredictionScore  # noqa: F401
