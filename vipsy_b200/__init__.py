"""vipsy_b200: B200-native ELBO + gradient path for vipsy's VI estimators.

    from vipsy_b200 import VIRT, VaeIRT, VCCDM, VCHoDina, Adam, Trace_ELBO
    VIRT(data=y, model="irt_2pl").fit(optim=Adam({"lr": 1e-2}), max_iter=5000)

Importing the package does not need a GPU; constructing an estimator does.
"""
from .infer import (Adam, MultiStepLR, PyroLRScheduler, StepLR, SVI, Trace_ELBO, TraceEnum_ELBO,  # noqa: F401
                    clear_param_store, constraints, get_param_store, param)
from .models import (BaseCDM, BaseIRT, BasePsy, BinEncoder, MvnEncoder, NormEncoder, SoftmaxEncoder,  # noqa: F401
                     VCCDM, VCDM, VCHoDina, VIRT, VaeCCDM, VaeCDM, VaeCHoDina, VaeIRT)

__all__ = ["VIRT", "VaeIRT", "VCDM", "VaeCDM", "VCCDM", "VaeCCDM", "VCHoDina", "VaeCHoDina",
           "Adam", "MultiStepLR", "StepLR", "PyroLRScheduler", "SVI", "Trace_ELBO", "TraceEnum_ELBO",
           "param", "clear_param_store", "get_param_store", "constraints"]
