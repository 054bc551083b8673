"""Frozen 3DGS map container (mirror of the part of gaussian_splatting/scene/gaussian_model.py the
tracking path uses: load_ply :238-327 and the five activated getters :75-95).  The reference's
densify / prune / optimiser utilities are mapping-time code that GS-EVT never calls; they are not here."""
import numpy as np
import torch
from torch import nn

from gsevt.compat import read_ply_vertices, write_ply_vertices


class GaussianModel:
    def __init__(self, sh_degree: int, config=None, device="cuda"):
        self.active_sh_degree = 0
        self.max_sh_degree = sh_degree
        self.device = device
        e = torch.empty(0)
        self._xyz = self._features_dc = self._features_rest = self._scaling = self._rotation = self._opacity = e
        self.config = config
        self._packed = None

    # activations exactly as the reference applies them on every access
    @property
    def get_scaling(self):
        return torch.exp(self._scaling)

    @property
    def get_rotation(self):
        return torch.nn.functional.normalize(self._rotation)

    @property
    def get_xyz(self):
        return self._xyz

    @property
    def get_features(self):
        return torch.cat((self._features_dc, self._features_rest), dim=1)

    @property
    def get_opacity(self):
        return torch.sigmoid(self._opacity)

    def load_ply(self, path):
        v = read_ply_vertices(path)
        n = v.shape[0]
        col = lambda name: np.asarray(v[name], dtype=np.float32)
        xyz = np.stack([col("x"), col("y"), col("z")], axis=1)
        opac = col("opacity")[:, None]
        dc = np.stack([col("f_dc_0"), col("f_dc_1"), col("f_dc_2")], axis=1)[:, :, None]            # (P,3,1)
        rest_names = sorted([p for p in v.dtype.names if p.startswith("f_rest_")], key=lambda s: int(s.split("_")[-1]))
        ncoef = (self.max_sh_degree + 1) ** 2
        assert len(rest_names) == 3 * ncoef - 3, "PLY SH degree does not match sh_degree"
        rest = np.stack([col(nm) for nm in rest_names], axis=1).reshape(n, 3, ncoef - 1) if rest_names else np.zeros((n, 3, 0), np.float32)
        sc = np.stack([col(nm) for nm in sorted([p for p in v.dtype.names if p.startswith("scale_")], key=lambda s: int(s.split("_")[-1]))], axis=1)
        rot = np.stack([col(nm) for nm in sorted([p for p in v.dtype.names if p.startswith("rot")], key=lambda s: int(s.split("_")[-1]))], axis=1)
        # frozen map: the reference wraps these in nn.Parameter(...) whose default requires_grad=True silently re-enables
        # gradients for every map tensor (gaussian_model.py:298-323); nothing consumes them, so they are really frozen here
        mk = lambda a: nn.Parameter(torch.tensor(a, dtype=torch.float, device=self.device).contiguous(), requires_grad=False)
        self._xyz = mk(xyz)
        self._features_dc = nn.Parameter(torch.tensor(dc, dtype=torch.float, device=self.device).transpose(1, 2).contiguous(), requires_grad=False)
        self._features_rest = nn.Parameter(torch.tensor(rest, dtype=torch.float, device=self.device).transpose(1, 2).contiguous(), requires_grad=False)
        self._opacity, self._scaling, self._rotation = mk(opac), mk(sc), mk(rot)
        self.active_sh_degree = self.max_sh_degree
        self._packed = None

    def save_ply(self, path):
        xyz = self._xyz.detach().cpu().numpy()
        f_dc = self._features_dc.detach().transpose(1, 2).flatten(start_dim=1).cpu().numpy()
        f_rest = self._features_rest.detach().transpose(1, 2).flatten(start_dim=1).cpu().numpy()
        names = ["x", "y", "z", "nx", "ny", "nz"] + [f"f_dc_{i}" for i in range(f_dc.shape[1])] + \
                [f"f_rest_{i}" for i in range(f_rest.shape[1])] + ["opacity"] + \
                [f"scale_{i}" for i in range(self._scaling.shape[1])] + [f"rot_{i}" for i in range(self._rotation.shape[1])]
        data = np.concatenate([xyz, np.zeros_like(xyz), f_dc, f_rest, self._opacity.detach().cpu().numpy(),
                               self._scaling.detach().cpu().numpy(), self._rotation.detach().cpu().numpy()], axis=1)
        arr = np.empty(xyz.shape[0], dtype=[(nm, "f4") for nm in names])
        for i, nm in enumerate(names):
            arr[nm] = data[:, i]
        write_ply_vertices(path, arr)

    def packed(self):
        """Engine-resident packed copy (built once: the map is frozen)."""
        if self._packed is None:
            from gsevt.engine import PackedMap
            self._packed = PackedMap.from_gaussian_model(self)
        return self._packed
