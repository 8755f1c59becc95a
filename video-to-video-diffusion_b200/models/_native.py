"""Shared plumbing for the nn.Module mirrors: keeps a libb2v object in sync with the module's parameters."""
import ctypes

import torch

from .. import _lib


def _weights_version(module):
    return sum(p._version for p in module.parameters()) + sum(b._version for b in module.buffers())


class NativeHandle:
    """Owns one b2v_unet / b2v_vae.  (Re)built lazily from the module's state_dict whenever a parameter changed
    (load_state_dict, .to(), optimiser step), so reference checkpoints load through the usual torch path."""

    def __init__(self, kind):
        self.kind = kind  # "unet" | "vae"
        self.handle = None
        self.version = None
        self.device = None

    def get(self, module, desc, device):
        ver = _weights_version(module)
        if self.handle is not None and self.version == ver and self.device == device:
            return self.handle
        self.close()
        L = _lib.lib()
        h = ctypes.c_void_p()
        with torch.cuda.device(device):
            _lib.check(getattr(L, f"b2v_{self.kind}_create")(ctypes.byref(h), ctypes.byref(desc)), f"{self.kind}_create")
            load = getattr(L, f"b2v_{self.kind}_load_weight")
            for key, val in module.state_dict().items():
                w = val.detach().to("cpu", torch.float32).contiguous()
                shape = (ctypes.c_int64 * max(1, w.dim()))(*w.shape)
                _lib.check(load(h, key.encode(), _lib.hptr(w), shape, w.dim()), f"load_weight({key})")
            _lib.check(getattr(L, f"b2v_{self.kind}_finalize")(h), f"{self.kind}_finalize")
        self.handle, self.version, self.device = h, ver, device
        return h

    def __deepcopy__(self, memo):  # copies of the module rebuild their own native object on first use
        return NativeHandle(self.kind)

    def __getstate__(self):
        return {"kind": self.kind, "handle": None, "version": None, "device": None}

    def close(self):
        if self.handle is not None:
            try:
                getattr(_lib.lib(), f"b2v_{self.kind}_destroy")(self.handle)
            except Exception:
                pass
            self.handle = None

    def __del__(self):
        self.close()


def require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(f"{what}: input is on {t.device}; this implementation runs on B200 only (no CPU fallback)")
    return t.detach().to(torch.float32).contiguous()


class ParamsOnly(torch.nn.Module):
    """Container that holds parameters under the reference's names; its math runs inside the parent's fused
    native call, so calling it directly is an error rather than a silent torch fallback."""

    def forward(self, *a, **k):
        raise RuntimeError(f"{type(self).__name__} is a parameter container; run the enclosing UNet3D / VAE module "
                           "(its forward executes on libb2v.so)")
