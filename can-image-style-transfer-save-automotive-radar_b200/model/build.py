"""Plugin factory with the reference's signature (IST/model/build.py:4-7)."""
from . import meta_arch


def build_model(cfg, pool='max'):
    model_factory = getattr(meta_arch, cfg.MODEL.META_ARCHITECTURE)
    model = model_factory(cfg, pool)
    return model
