"""Dry-run the GPU test modules on the CPU with the oracle standing in for the device arena: validates the
test code itself (API names, shapes, scene logic) before spending GPU minutes.  Not part of the test suite."""
import sys, os, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # repo root (this file lives in tests/)
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import orc
import mars_ode_physics_b200 as pkg
from mars_ode_physics_b200 import capi

class FakeArena(orc.OracleBatch):
    def __init__(self, W, device=0, strict=False):
        super().__init__(W, single=False)
        self.h = None
        self._launches = 0
        self._kernel = 1
    def lcp_residual(self):
        return np.array([self.last_residual(w) for w in range(self.num_worlds)], dtype=np.float32)
    def set_sor_kernel(self, k): self._kernel = k
    def sor_kernel_used(self): return self._kernel
    def set_sor_warps(self, n): pass
    def set_tune(self, bits): pass
    def set_stepper(self, s): self._stepper = s; super().set_stepper(s)
    def stepper(self): return getattr(self, '_stepper', 0)
    def dense_failures(self): return sum(self.last_lcp_error(w) for w in range(self.num_worlds))
    def dense_residual(self): return np.array([self.last_residual(w) for w in range(self.num_worlds)], dtype=np.float32)
    def set_register_rows(self, n): pass
    def set_resident_rows(self, n): pass
    def set_graph_launch(self, on): pass
    def collide_record_pairs(self, on=True): pass
    def contacts_dropped(self): return 0
    def kernel_launch_count(self): return self._launches
    def collide(self):
        self._launches += 1; super().collide()
    def step(self, h):
        self._launches += 3; super().step(h)
    def collide_step(self, h):
        self.collide(); self.step(h)
    def state_upload(self, st):
        nb = st.shape[0]
        for w in range(self.num_worlds):
            for b in range(nb):
                self.body_set_state(b, st[b, :, w].astype(np.float64), w)
    def state_download(self, nb, out=None):
        return super().state_download(nb).astype(np.float32)
    def body_get_force(self, b, world=0):
        return super().body_get_force(b, world)

pkg.Arena = FakeArena
import pytest
import conftest
os.environ["B2P_FAKE_ARENA"] = "1"
mods = sys.argv[1:] or ["tests/test_gpu_epilogue.py"]
sys.exit(pytest.main(["-x", "-q", "-m", "gpu", "-p", "no:cacheprovider", *mods, "-k", os.environ.get("K", "")]))
