"""Field shims: the small part of the Taichi field API that the reference's callers use on
ParticleSystem / solver objects (main.py:76,111,159-161,170,173,190,196): `x[None]`, `x[i]`,
`.to_numpy()`, `.from_numpy()`, `.fill()`, `.shape`.  Storage is a torch CUDA tensor (or a view of
one); nothing here computes on the CPU beyond host copies the caller asks for.
"""
import numpy as np
import torch


class TensorField:
    """A field backed by a device tensor (possibly a strided view such as pos4[:, :3])."""

    def __init__(self, tensor):
        self._t = tensor

    @property
    def tensor(self):
        return self._t

    @property
    def shape(self):
        return tuple(self._t.shape[:1]) if self._t.dim() > 0 else ()

    def to_numpy(self):
        return self._t.detach().contiguous().cpu().numpy()

    def to_torch(self, device=None):
        t = self._t.detach().clone()
        return t.to(device) if device is not None else t

    def from_numpy(self, arr):
        a = torch.from_numpy(np.ascontiguousarray(arr)).to(self._t.dtype)
        self._t.copy_(a.reshape(self._t.shape).to(self._t.device))

    def fill(self, value):
        if isinstance(value, (int, float)):
            self._t.fill_(value)
        else:
            v = torch.as_tensor(np.asarray(value, dtype=np.float64), dtype=self._t.dtype, device=self._t.device)
            self._t.copy_(v.expand_as(self._t))

    def __getitem__(self, idx):
        if idx is None:
            v = self._t.detach().cpu().numpy()
            return v.item() if v.ndim == 0 else v
        v = self._t[idx].detach().cpu().numpy()
        return v.item() if v.ndim == 0 else v

    def __setitem__(self, idx, value):
        v = torch.as_tensor(np.asarray(value), dtype=self._t.dtype, device=self._t.device)
        if idx is None:
            self._t.copy_(v.reshape(self._t.shape))
        else:
            self._t[idx] = v

    def __len__(self):
        return self._t.shape[0]


class HostScalar:
    """0-d field whose value lives on the host (e.g. exist_rigid, active_rigid, simulate_cnt)."""

    def __init__(self, value=0, on_set=None):
        self._v = value
        self._on_set = on_set

    def __getitem__(self, idx):
        return self._v

    def __setitem__(self, idx, value):
        self._v = value
        if self._on_set is not None:
            self._on_set(value)

    def to_numpy(self):
        return np.asarray(self._v)


class DeviceScalar:
    """0-d field whose authoritative value lives in the library's device control block (delta_time):
    reading it synchronises exactly like `field[None]` does in the reference."""

    def __init__(self, getter, setter):
        self._get = getter
        self._set = setter

    def __getitem__(self, idx):
        return self._get()

    def __setitem__(self, idx, value):
        self._set(value)

    def to_numpy(self):
        return np.asarray(self._get())


class FetchedField:
    """Per-particle result that lives in the library's sorted scratch; `.to_numpy()` fetches it back
    in original particle order (sph_fetch)."""

    def __init__(self, ps, field_id, ncomp=1, dtype=torch.float32):
        self._ps = ps
        self._id = field_id
        self._nc = ncomp
        self._dtype = dtype

    @property
    def shape(self):
        return (self._ps.particle_num,)

    def to_torch(self):
        width = 4 if self._nc == 3 else 1
        t = self._ps._fetch(self._id, width, self._dtype)
        return t[:, :3] if self._nc == 3 else t.reshape(-1)

    def to_numpy(self):
        return self.to_torch().contiguous().cpu().numpy()

    def __getitem__(self, idx):
        v = self.to_numpy()
        return v if idx is None else v[idx]


class ParticleFields:
    """Struct-of-fields view (the reference's `Particles.field(shape=n)`, PS:6-20), SoA underneath."""

    def __init__(self, **fields):
        for k, v in fields.items():
            setattr(self, k, v)
