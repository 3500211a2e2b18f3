"""TEST INFRASTRUCTURE — ctypes front end of the compiled REFERENCE rasterizer (oracle/_ref/
libref_*.so, built by oracle/build.py from the unmodified sources).  Needs a GPU.  Only tests/,
tests/golden/make_golden.py and bench.py's reference-CUDA comparison import this module."""
from __future__ import annotations

import ctypes as C
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
_libs = {}

FIELDS = {  # name -> (dtype, trailing shape or None for flat)
    "depths": (torch.float32, ()), "clamped": (torch.uint8, (3,)), "means2D": (torch.float32, (2,)),
    "cov3D": (torch.float32, (6,)), "conic_opacity": (torch.float32, (4,)), "rgb": (torch.float32, (3,)),
    "tiles_touched": (torch.int32, ()), "point_offsets": (torch.int32, ()),
}


def available(name: str) -> bool:
    return os.path.exists(os.path.join(REF_DIR, f"libref_{name}.so"))


def load(name: str):
    if name not in _libs:
        path = os.path.join(REF_DIR, f"libref_{name}.so")
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path}: run `python oracle/build.py` where /root/reference exists")
        L = C.CDLL(path, mode=os.RTLD_LOCAL)
        L.ref_last_error.restype = C.c_char_p
        L.ref_state_new.restype = C.c_void_p
        L.ref_state_free.argtypes = [C.c_void_p]
        L.ref_state_field.restype = C.c_longlong
        L.ref_state_field.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
        vp, i, f = C.c_void_p, C.c_int, C.c_float
        L.ref_forward.argtypes = [vp, i, i, i, vp, i, i, vp, vp, vp, vp, vp, f, vp, vp, vp, vp, vp, f, f, i, i, vp, vp, vp, i]
        L.ref_backward.argtypes = [vp, i, i, i, i, vp, i, i, vp, vp, vp, vp, f, vp, vp, vp, vp, vp, f, f, vp,
                                   vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i]
        L.ref_mark_visible.argtypes = [i, vp, vp, vp, vp]
        _libs[name] = L
    return _libs[name]


def _p(t):
    return None if t is None else t.data_ptr()


class RefRasterizer:
    """One forward/backward pair of the reference library on CUDA tensors (float32, contiguous)."""

    def __init__(self, name: str):
        self.name = name
        self.L = load(name)
        self.state = self.L.ref_state_new()
        self.is_rgbd = bool(self.L.ref_is_rgbd())
        self.bwd_channels = int(self.L.ref_num_channels_bwd())

    def __del__(self):
        try:
            self.L.ref_state_free(self.state)
        except Exception:
            pass

    def forward(self, *, bg, means3D, opacities, viewmatrix, projmatrix, campos, tanfovx, tanfovy, W, H,
                shs=None, colors_precomp=None, scales=None, rotations=None, cov3D_precomp=None,
                scale_modifier=1.0, sh_degree=3, num_channels=3, debug=False):
        dev = means3D.device
        P = means3D.shape[0]
        M = shs.shape[1] if shs is not None else 0
        self.args = dict(bg=bg, means3D=means3D, shs=shs, colors_precomp=colors_precomp, scales=scales,
                         rotations=rotations, cov3D_precomp=cov3D_precomp, viewmatrix=viewmatrix,
                         projmatrix=projmatrix, campos=campos, tanfovx=tanfovx, tanfovy=tanfovy, W=W, H=H, P=P,
                         M=M, D=sh_degree, scale_modifier=scale_modifier, C=num_channels)
        color = torch.zeros((num_channels, H, W), dtype=torch.float32, device=dev)
        depth = torch.zeros((1, H, W), dtype=torch.float32, device=dev) if self.is_rgbd else None
        radii = torch.zeros((P,), dtype=torch.int32, device=dev)
        R = self.L.ref_forward(self.state, P, sh_degree, M, _p(bg), W, H, _p(means3D), _p(shs), _p(colors_precomp),
                               _p(opacities), _p(scales), scale_modifier, _p(rotations), _p(cov3D_precomp),
                               _p(viewmatrix), _p(projmatrix), _p(campos), tanfovx, tanfovy, 0, num_channels,
                               _p(color), _p(depth), _p(radii), int(debug))
        if R < 0:
            raise RuntimeError("reference forward failed: " + self.L.ref_last_error().decode())
        torch.cuda.synchronize(dev)
        self.R, self.radii = R, radii
        return dict(R=R, color=color, depth=depth, radii=radii)

    def field(self, name: str) -> torch.Tensor:
        a = self.args
        dev = a["means3D"].device
        P, W, H = a["P"], a["W"], a["H"]
        tiles = ((W + 15) // 16) * ((H + 15) // 16)
        if name in FIELDS:
            dt, tail = FIELDS[name]
            t = torch.zeros((P,) + tail, dtype=dt, device=dev)
        elif name == "accum_alpha":
            t = torch.zeros((H * W,), dtype=torch.float32, device=dev)
        elif name == "n_contrib":
            t = torch.zeros((H * W,), dtype=torch.int32, device=dev)
        elif name == "ranges":
            t = torch.zeros((tiles, 2), dtype=torch.int32, device=dev)
        elif name == "point_list":
            t = torch.zeros((max(self.R, 1),), dtype=torch.int32, device=dev)
        elif name == "point_list_keys":
            t = torch.zeros((max(self.R, 1),), dtype=torch.int64, device=dev)
        else:
            raise KeyError(name)
        n = self.L.ref_state_field(self.state, name.encode(), t.data_ptr())
        if n < 0:
            raise RuntimeError(f"reference state field {name} unavailable")
        torch.cuda.synchronize(dev)
        if name in ("point_list", "point_list_keys"):
            t = t[: self.R]
        return t

    def backward(self, dL_dpix, debug=False):
        a = self.args
        dev = a["means3D"].device
        P, M = a["P"], a["M"]
        Cb = self.bwd_channels
        if dL_dpix.shape[0] != Cb:
            raise ValueError(f"this reference build differentiates exactly {Cb} channels")
        z = lambda *s: torch.zeros(s, dtype=torch.float32, device=dev)
        g = dict(dL_dmeans2D=z(P, 3), dL_dconic=z(P, 4), dL_dopacity=z(P), dL_dcolors=z(P, Cb), dL_dmeans3D=z(P, 3),
                 dL_dcov3D=z(P, 6), dL_dsh=z(P, max(M, 1), 3), dL_dscales=z(P, 3), dL_drotations=z(P, 4))
        rc = self.L.ref_backward(self.state, P, a["D"], M, self.R, _p(a["bg"]), a["W"], a["H"], _p(a["means3D"]),
                                 _p(a["shs"]), _p(a["colors_precomp"]), _p(a["scales"]), a["scale_modifier"],
                                 _p(a["rotations"]), _p(a["cov3D_precomp"]), _p(a["viewmatrix"]), _p(a["projmatrix"]),
                                 _p(a["campos"]), a["tanfovx"], a["tanfovy"], _p(self.radii), _p(dL_dpix.contiguous()),
                                 _p(g["dL_dmeans2D"]), _p(g["dL_dconic"]), _p(g["dL_dopacity"]), _p(g["dL_dcolors"]),
                                 _p(g["dL_dmeans3D"]), _p(g["dL_dcov3D"]), _p(g["dL_dsh"]), _p(g["dL_dscales"]),
                                 _p(g["dL_drotations"]), int(debug))
        if rc != 0:
            raise RuntimeError("reference backward failed: " + self.L.ref_last_error().decode())
        torch.cuda.synchronize(dev)
        g["dL_dsh"] = g["dL_dsh"][:, :M]
        return g
