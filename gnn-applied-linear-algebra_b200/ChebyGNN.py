"""Drop-in for pytorch/ChebyGNN.py: Chebyshev relaxation of degree `deg`."""
import torch

from . import _runtime as rt
from ._io import Placement, float_dtype


def _recurrence(deg, g):
    """The scalar recurrences of the reference's global updates, evaluated with the same torch
    expressions on 0-d tensors of g's dtype/device (ChebyGNN.py:137, :262-263, :282-283), so the
    scalars are bit-identical.  Returns per-iteration (alpha_old, alpha, beta) and the final g."""
    c, d = g[0], g[1]
    rows = []
    alpha = 1 / d
    beta = torch.zeros_like(alpha)
    rows.append((alpha, alpha, beta))
    g_out = torch.hstack([c, d, alpha])
    for it in range(2, deg + 1):
        alpha_old = alpha
        beta = 0.5 * (c * alpha) ** 2 if it == 2 else ((c * alpha) / 2) ** 2
        alpha = 1 / (d - beta / alpha)
        rows.append((alpha_old, alpha, beta))
        g_out = torch.hstack([c, d, alpha, beta])
    return rows, g_out


_table_cache = {}


def _device_table(deg, g, device, dt):
    """(per-iteration [alpha_old, alpha, beta] table on the device, final g) for host-resident g, cached
    per (deg, c, d, dtypes, device): the recurrence is ~20 tiny CPU tensor ops plus an upload, more host
    time per call than the kernels take on a partitioned operator."""
    if g.is_cuda:
        rows, g_out = _recurrence(deg, g)
        return torch.stack([torch.stack(r) for r in rows]).to(device=device, dtype=dt).contiguous(), g_out
    key = (deg, float(g[0]), float(g[1]), g.dtype, str(device), dt)
    hit = _table_cache.get(key)
    if hit is None:
        if len(_table_cache) > 64:
            _table_cache.clear()
        rows, g_out = _recurrence(deg, g)
        hit = (torch.stack([torch.stack(r) for r in rows]).to(device=device, dtype=dt).contiguous(), g_out)
        _table_cache[key] = hit
    return hit[0], hit[1].clone()


class ChebyRelaxGNN(torch.nn.Module):
    """ChebyGNN.py:287-353.  in: vertex_attr=[b,x], edge_attr=[A_ij], g=[c,d];
    out: vertex_attr=[b,x,r,p], edge_attr=[A_ij,z_ij], g=[c,d,alpha,beta].
    Iteration 1 is one launch of glab_cheby_first, every later iteration one launch of
    glab_cheby_next (SpMV + both vertex updates fused; the reference runs 2 GN blocks each).
    Extension: vertex_attr = [b (k cols) | x (k cols)] relaxes k right-hand sides."""

    def __init__(self, deg=3):
        super().__init__()
        self.deg = deg

    def _forward_partitioned(self, vertex_attr, pg, edge_attr, g):
        """This rank's row block of a row-partitioned operator (edgeij_pair = dist.PartitionedGraph)."""
        io = Placement(vertex_attr, edge_attr)
        dt = float_dtype(vertex_attr, edge_attr)
        n, F = vertex_attr.shape
        k = F // 2
        op = pg.operator(edge_attr, k, dt)
        va = io.up(vertex_attr, dt)
        ent = op.entry()
        b, _ = rt.unpack(va, [(0, k), (k, k)], outs=[None, op.local(ent)])
        op.publish(ent)
        table, g_out = _device_table(self.deg, g, io.device, dt)
        x, r, pname = op.chebyshev(self.deg, b, table, ent)
        if op.halo.part.world > 1:
            op.acquire(op.last_gathered)      # the message column reads that vector's halo tail
        e_out = rt.with_messages(op.plan, op.vals, op.vec[op.last_gathered])
        v_out = rt.pack([b, x, r, op.local(pname)])
        return io.down(v_out), io.down(e_out), g_out

    def forward(self, vertex_attr, edgeij_pair, edge_attr, g, batch=None):
        if self.deg <= 0:
            return vertex_attr, edge_attr, g
        from .dist import is_partitioned
        if is_partitioned(edgeij_pair):
            return self._forward_partitioned(vertex_attr, edgeij_pair, edge_attr, g)
        io = Placement(vertex_attr, edgeij_pair, edge_attr)
        dt = float_dtype(vertex_attr, edge_attr)
        n, F = vertex_attr.shape
        k = F // 2
        plan = rt.get_plan(edgeij_pair, n)
        vals = rt.get_vals(plan, edge_attr, 0, dt)
        va = io.up(vertex_attr, dt)
        b, x0 = rt.unpack(va, [(0, k), (k, k)])
        table, g_out = _device_table(self.deg, g, io.device, dt)
        x = torch.empty_like(x0)
        r = torch.empty_like(x0)
        p = torch.empty_like(x0)
        p_alt = torch.empty_like(x0)
        rt.cheby_first(plan, vals, b, x0, x, r, p, table[0, 1:2])
        gathered = x0  # vector whose A_ij * v_j messages the last SpMV block produced
        for it in range(1, self.deg):
            rt.cheby_next(plan, vals, p, p_alt, r, x, table[it, 0:1], table[it, 1:2], table[it, 2:3])
            gathered = p
            p, p_alt = p_alt, p
        e_out = rt.with_messages(plan, vals, gathered)
        v_out = rt.pack([b, x, r, p])
        return io.down(v_out), io.down(e_out), g_out
