"""go-vectorsearch_b200 -- B200-native (sm_100a) backend for go-vectorsearch's similarity-search hot path.

Host-side mirror of the reference's interfaces over the C ABI in include/vscuda.h:

  compute  -- compute/types.go, compute/compute.go, compute/cosine.go, compute/quantization.go
  ivf      -- server/search.go:202-273 (device-resident IVF-Flat probe-and-scan)
  dnc      -- dnc/k_means.go:67-117 (one Lloyd step), dnc/dnc.go:417-449 (recenter)

There is no CPU fallback: importing works anywhere, but every compute call needs libvscuda.so and a
B200; both failures raise loudly (BackendUnavailable).
"""
from . import _lib  # noqa: F401
from ._lib import BackendUnavailable, load, lib_path  # noqa: F401
from . import compute, ivf, dnc, shard  # noqa: F401
