"""spectro_b200 — host layer of the B200-native spectrogram engine (drop-in for the render
path of triq-org/spectroplot-js: Spectroplot options / setData API and the worker message
protocol above a C ABI; all arithmetic of the path runs in hand-written sm_100a kernels)."""
from ._lib import Engine, PinnedBuffer, SpError, FORMATS  # noqa: F401
from .worker import GpuWorker, renderFft  # noqa: F401
from .spectroplot import Spectroplot  # noqa: F401
from . import windows, cmaps, sharding, ingest, egress  # noqa: F401
