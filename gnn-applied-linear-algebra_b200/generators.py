"""Operator generators (inputs of the hot path, not the hot path): constant-stencil matrices
on an N x N grid as the reference's COO layout -- ``edge_index`` int64 [2, z] row-major sorted
with the diagonal included, ``edge_val`` [z, 1] -- built with vectorised torch ops on any device.

  laplacian_2d            UtilsGNN.py:53-67 (kron(I,T)+kron(T,I), T = tridiag(1,-2,1))
  heat_fem_2d             matlab/heateqnfem2dfun.m:52-172 with bcs=[2,2] (eliminated Dirichlet)
  constant_diffusion_fem  DiffCoeffs/FEM.py:184-198 (periodic Q1 FEM, D = diag(alpha, beta))
"""
import torch

_OFFS = [(-1, -1), (-1, 0), (-1, 1), (0, -1), (0, 0), (0, 1), (1, -1), (1, 0), (1, 1)]  # (dy, dx)


def stencil_coo(ny, nx, weights, periodic=False, dtype=torch.float64, device="cpu", rows=None):
    """weights: {(dy, dx): value}; grid point (y, x) has index y*nx + x (x fastest).
    Dirichlet truncation drops out-of-grid neighbours; periodic wraps them.  Columns are
    emitted in ascending order within each row (== scipy/torch coalesced COO order)."""
    device = torch.device(device)
    if rows is None:
        rows = (0, ny * nx)
    idx = torch.arange(rows[0], rows[1], dtype=torch.int64, device=device)   # GLOBAL row ids of this block
    n = idx.numel()
    gy, gx = idx // nx, idx % nx
    offs = [o for o in _OFFS if o in weights and weights[o] != 0]
    cols, keep = [], []
    for dy, dx in offs:
        yy, xx = gy + dy, gx + dx
        if periodic:
            cols.append((yy % ny) * nx + (xx % nx))
            keep.append(torch.ones(n, dtype=torch.bool, device=device))
        else:
            cols.append(yy * nx + xx)
            keep.append((yy >= 0) & (yy < ny) & (xx >= 0) & (xx < nx))
    cols = torch.stack(cols, 1)
    keep = torch.stack(keep, 1)
    vals = torch.tensor([weights[o] for o in offs], dtype=dtype, device=device).expand(n, len(offs))
    if periodic:
        # wrap-around breaks the ascending column order inside a row: sort each row's entries
        cols, order = torch.sort(cols, dim=1, stable=True)
        vals = torch.gather(vals, 1, order)
    ridx = idx.view(-1, 1).expand(n, len(offs))
    edge_index = torch.stack([ridx[keep], cols[keep]])
    edge_val = vals[keep].reshape(-1, 1).contiguous()
    return edge_index, edge_val


def laplacian_2d(N, dtype=torch.float64, device="cpu", rows=None):
    """5-point (negative) Laplacian, diag -4, off-diag +1, Dirichlet truncation.
    rows=(r0, r1) builds only that block of rows (global row / column ids), for the
    row-partitioned multi-GPU path."""
    w = {(0, 0): -4.0, (0, -1): 1.0, (0, 1): 1.0, (-1, 0): 1.0, (1, 0): 1.0}
    return stencil_coo(N, N, w, False, dtype, device, rows)


def heat_fem_stencil(hx=1.0, hy=1.0):
    """Interior 9-point stencil of heateqnfem2dfun.m:91 (element values) assembled over the four
    elements that share a node: centre 4*e0, E/W 2*e1, N/S 2*e2, corners e3."""
    a = hy / hx
    e = [(2 * a * a + 2) / (6 * a), (-2 * a * a + 1) / (6 * a), (a * a - 2) / (6 * a), (-1 - a * a) / (6 * a)]
    w = {(0, 0): 4 * e[0], (0, -1): 2 * e[1], (0, 1): 2 * e[1], (-1, 0): 2 * e[2], (1, 0): 2 * e[2]}
    for o in ((-1, -1), (-1, 1), (1, -1), (1, 1)):
        w[o] = e[3]
    return w


def heat_fem_2d(num_cells, h=(1.0, 1.0), dtype=torch.float64, device="cpu", rows=None):
    """heateqnfem2dfun(num_cells, h, [2,2]): (num_cells-1) interior nodes per direction."""
    mx, my = num_cells[0] - 1, num_cells[1] - 1
    return stencil_coo(my, mx, heat_fem_stencil(h[0], h[1]), False, dtype, device, rows)


def constant_diffusion_stencil(alpha, beta):
    return {(0, 0): (4.0 / 3.0) * (alpha + beta),
            (0, -1): (-2 * alpha + beta) / 3.0, (0, 1): (-2 * alpha + beta) / 3.0,
            (-1, 0): (alpha - 2 * beta) / 3.0, (1, 0): (alpha - 2 * beta) / 3.0,
            (-1, -1): -(alpha + beta) / 6.0, (-1, 1): -(alpha + beta) / 6.0,
            (1, -1): -(alpha + beta) / 6.0, (1, 1): -(alpha + beta) / 6.0}


def constant_diffusion_fem(alpha, beta, N, dtype=torch.float64, device="cpu"):
    """ConstantDiffusionFEM_Builder().generate_problem_stiffness_matrix(alpha, beta, N),
    coalesced: periodic 9-point operator on the N x N torus."""
    return stencil_coo(N, N, constant_diffusion_stencil(alpha, beta), True, dtype, device)


def diagonal_of(edge_index, edge_val, n):
    """A_ii as a dense [n, 1] tensor (duplicates summed)."""
    on = edge_index[0] == edge_index[1]
    d = torch.zeros(n, dtype=edge_val.dtype, device=edge_val.device)
    d.index_add_(0, edge_index[0][on], edge_val.reshape(-1)[on])
    return d.view(-1, 1)
