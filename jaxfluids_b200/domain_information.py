"""DomainInformation for homogeneous meshes and block decomposition.

Mirrors the parts of domain/domain_information.py the path uses: cell sizes
(mesh_creation/homogenous.py:7-20), 1/dx (:290), smallest cell size (:697-702),
interior slices (:domain_slices_conservatives), and the decomposition
bookkeeping split_x*split_y*split_z with block (i,j,k) <-> rank i*sy*sz + j*sz + k
(domain/helper_functions.py:155-169, :63-80).
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

AXES = ("x", "y", "z")
FACES = ("east", "west", "north", "south", "top", "bottom")


class DomainInformation:
    face_location_to_axis_index = {"east": 0, "west": 0, "north": 1, "south": 1, "top": 2, "bottom": 2}

    def __init__(self, cells, domain_range, split, nh: int):
        self.global_number_of_cells = tuple(int(c) for c in cells)
        self.global_domain_size = tuple((float(a), float(b)) for a, b in domain_range)
        self.split_factors = tuple(int(s) for s in split)
        self.nh_conservatives = int(nh)
        self.no_subdomains = int(np.prod(self.split_factors))
        self.is_parallel = self.no_subdomains > 1
        for ax, n, sp in zip(AXES, self.global_number_of_cells, self.split_factors):
            assert sp >= 1 and n % sp == 0, (
                f"domain/decomposition/split_{ax}={sp} does not divide domain/{ax}/cells={n}")
        self.device_number_of_cells = tuple(n // s for n, s in zip(self.global_number_of_cells, self.split_factors))
        self.active_axes_indices = tuple(i for i in range(3) if self.global_number_of_cells[i] > 1)
        self.inactive_axes_indices = tuple(i for i in range(3) if self.global_number_of_cells[i] == 1)
        self.active_axes = tuple(AXES[i] for i in self.active_axes_indices)
        self.inactive_axes = tuple(AXES[i] for i in self.inactive_axes_indices)
        self.dim = len(self.active_axes_indices)
        self.active_face_locations = tuple(f for f in FACES if self.face_location_to_axis_index[f] in self.active_axes_indices)
        nh = self.nh_conservatives
        self.domain_slices_conservatives = tuple(
            slice(nh, -nh) if i in self.active_axes_indices else slice(None) for i in range(3))
        # homogenous.py:13 ; domain_information.py:290
        self.cell_sizes = tuple(np.float64((hi - lo) / n) for (lo, hi), n in
                                zip(self.global_domain_size, self.global_number_of_cells))
        self.one_cell_sizes = tuple(np.float64(1.0) / d for d in self.cell_sizes)
        self.smallest_cell_size = float(min(self.cell_sizes[i] for i in self.active_axes_indices))

    # -- mesh -----------------------------------------------------------------
    def get_global_cell_centers(self):
        out = []
        for (lo, hi), n in zip(self.global_domain_size, self.global_number_of_cells):
            d = (hi - lo) / n
            out.append(np.linspace(lo + d / 2, hi - d / 2, n))      # homogenous.py:14
        return tuple(out)

    def get_device_cell_centers(self, rank: int = 0):
        idx = self.block_index(rank)
        cc = self.get_global_cell_centers()
        return tuple(c[i * m:(i + 1) * m] for c, i, m in zip(cc, idx, self.device_number_of_cells))

    def compute_device_mesh_grid(self, rank: int = 0, sparse: bool = False):
        """domain_information.py:363-375. sparse=True returns broadcastable axes (same values)."""
        cc = self.get_device_cell_centers(rank)
        mg = np.meshgrid(*cc, indexing="ij", sparse=sparse)
        return tuple(mg[i] for i in self.active_axes_indices)

    def compute_global_mesh_grid(self):
        mg = np.meshgrid(*self.get_global_cell_centers(), indexing="ij")
        return tuple(mg[i] for i in self.active_axes_indices)

    # -- shapes ------------------------------------------------------------------
    @property
    def device_shape_with_halos(self):
        nh = self.nh_conservatives
        return (5,) + tuple(n + 2 * nh if N > 1 else 1 for n, N in
                            zip(self.device_number_of_cells, self.global_number_of_cells))

    @property
    def cells_per_device(self) -> int:
        return int(np.prod(self.device_number_of_cells))

    # -- decomposition -------------------------------------------------------------
    def block_index(self, rank: int) -> Tuple[int, int, int]:
        sx, sy, sz = self.split_factors
        return (rank // (sy * sz), (rank // sz) % sy, rank % sz)

    def rank_of(self, idx) -> int:
        sx, sy, sz = self.split_factors
        return (idx[0] % sx) * sy * sz + (idx[1] % sy) * sz + (idx[2] % sz)

    def block_slices(self, rank: int):
        """Interior slices of block `rank` in a global interior-only array (split_buffer_np)."""
        idx = self.block_index(rank)
        return tuple(slice(i * m, (i + 1) * m) for i, m in zip(idx, self.device_number_of_cells))

    def neighbor(self, rank: int, face: str, periodic: bool):
        """Rank owning the block across `face`, or None at a physical (non-periodic) boundary.
        halos/inner/halo_communication.py:35-77 (wrapped permutation), :103-150 (mask)."""
        ax = self.face_location_to_axis_index[face]
        step = 1 if face in ("east", "north", "top") else -1
        idx = list(self.block_index(rank))
        s = self.split_factors[ax]
        if s == 1:
            return None
        j = idx[ax] + step
        if j < 0 or j >= s:
            if not periodic:
                return None
            j %= s
        idx[ax] = j
        return self.rank_of(idx)
