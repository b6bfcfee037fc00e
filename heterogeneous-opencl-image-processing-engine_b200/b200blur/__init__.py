"""b200blur -- Python front end (ctypes) of the B200-native Gaussian-blur stream engine.

The product is ``libb200blur.so`` (C ABI in ``include/b200blur.h`` over hand-written sm_100a CUDA); this package
only binds it so tests and ``bench.py`` can drive it, mirroring the reference's host-side vocabulary
(``heterogeneous_blur.c`` / ``split_image_blur.c``): contexts, queues, write / blur / read enqueues, finish,
profiling events, per-batch distribution.  There is no CPU fallback: every compute call goes to the GPU library and
raises ``BlurError`` if it cannot.
"""
from .lib import (  # noqa: F401
    BlurError,
    Context,
    Feed,
    Launch,
    Stats,
    build,
    device_count,
    device_name,
    lib_path,
    load,
    partition,
    plan_groups,
    plan_row_edge,
    ratio_split_images,
    ratio_split_row,
    run_host_multi,
    version,
    DECLARED_SYMBOLS,
)
from .sharding import band_rows, image_shard, BandPlan, plan_bands  # noqa: F401
