from fsd_b200.sahi_api.postprocess import (GreedyNMMPostprocess, LSNMSPostprocess, NMMPostprocess, NMSPostprocess,  # noqa: F401
                                           PostprocessPredictions)
