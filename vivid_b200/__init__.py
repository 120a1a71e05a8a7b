"""vivid_b200 — B200-native (sm_100a) implementation of VIVID's guided EDM2 denoising hot path.

Public surface mirrors the reference's Python API for this path:
  NVPrecond                      the persisted net / gnet / sr_model object
  edm_sampler                    the sampler call of generate_images.py
  generate_images_nvs            the generation driver used by calculate_metrics.py gen
  calculate_stats_for_iterable_nvs, calculate_metrics_from_stats_nvs, get_metrics
                                 the statistics leg of calculate_metrics.py gen (detector networks are user-supplied)
  StandardRGBEncoder, StackedRandomGenerator, compose_geometry
All arithmetic runs in libvividb200.so (hand-written CUDA behind the C ABI of include/vivid_b200.h).
"""
from .precond import NVPrecond  # noqa: F401
from .sampler import StackedRandomGenerator, edm_sampler  # noqa: F401
from .encoders import StandardRGBEncoder  # noqa: F401
from .synthetic import compose_geometry  # noqa: F401
from .generate import generate_images_nvs, get_metrics  # noqa: F401
from .metrics import calculate_metrics_from_stats_nvs, calculate_stats_for_iterable_nvs  # noqa: F401

__all__ = ["NVPrecond", "edm_sampler", "StackedRandomGenerator", "StandardRGBEncoder", "compose_geometry",
           "generate_images_nvs", "get_metrics", "calculate_stats_for_iterable_nvs", "calculate_metrics_from_stats_nvs"]
