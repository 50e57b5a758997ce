"""Dry run of bench.py's OWN arm without a GPU (test infrastructure; used by tests/test_bench_cpu.py): torch.cuda's streams / events and the
product library are replaced by stand-ins and the CPU-EMULATED kernel library, then bench.main() runs unchanged on a small grid.  It checks
the Python of the measurement path (argument handling, byte models, JSON line) -- the numbers it prints mean nothing."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                     # noqa: E402
import impec2d_b200 as I         # noqa: E402
from impec2d_b200 import _lib    # noqa: E402


class FakeStream:
    cuda_stream = 0


class FakeEvent:
    def __init__(self, enable_timing=False): self.t = None
    def record(self, stream=None): self.t = time.perf_counter()
    def elapsed_time(self, other): return (other.t - self.t) * 1e3


torch.cuda.is_available = lambda: True
torch.cuda.set_device = lambda d: None
torch.cuda.Stream = FakeStream
torch.cuda.set_stream = lambda s: None
torch.cuda.Event = FakeEvent
torch.cuda.synchronize = lambda *a: None
_empty = torch.empty
torch.empty = lambda *a, **k: _empty(*a, **{kk: v for kk, v in k.items() if kk != "pin_memory"})

_load = _lib.load
_lib.load = lambda emulated=False: _load(True)                     # every library call goes to the emulated build
_default_params = I.default_params
I.default_params = lambda **kw: _default_params(**dict(kw, emulated=True))
I.api.default_params = I.default_params
_Context = I.Context


class EmuContext(_Context):
    def __init__(self, params, device=0, rank=0, nranks=1, nccl_id=None, stream=None, emulated=False):
        super().__init__(params, device, rank, nranks, nccl_id, None, emulated=True)

    def profile_kernel(self, name, reps=10):
        return max(super().profile_kernel(name, 1), 1e-3)         # the emulator's events do not time: any positive number will do


I.Context = EmuContext
for name in ("synth_state_rows", "init_source_rows"):
    fn = getattr(I, name)
    setattr(I, name, (lambda f: lambda *a, **k: f(*a, **dict(k, emulated=True)))(fn))


# ---- several ranks (torchrun, WORLD_SIZE > 1): gloo instead of NCCL, "cuda" tensors on the CPU, the emulated build's file communicator
# (IMPEC_EMU_COMM_DIR, csrc/emu_filecomm.inl) in place of the library's NCCL / peer-memory transport
if int(os.environ.get("WORLD_SIZE", "1")) > 1:
    import torch.distributed as dist
    _init = dist.init_process_group
    dist.init_process_group = lambda backend=None, **k: _init("gloo")
    _tensor = torch.tensor
    torch.tensor = lambda *a, **k: _tensor(*a, **{kk: v for kk, v in k.items() if kk != "device"})
    torch.device = (lambda f: lambda *a, **k: f("cpu"))(torch.device)

    class LibProxy:
        def __init__(self, lib): self._lib = lib
        def __getattr__(self, name): return getattr(self._lib, name)
        def impec_nccl_unique_id(self, buf): return 0               # the emulated build has no NCCL: the id is not used

    I.load = lambda emulated=False: LibProxy(_load(True))

import bench                      # noqa: E402
sys.exit(bench.main())
