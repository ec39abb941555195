"""Drop-in for pytorch/PowerMethodGNN.py: power iteration + Rayleigh quotient."""
import torch

from . import _runtime as rt
from ._io import Placement, float_dtype


class PowerMethodGNN(torch.nn.Module):
    """PowerMethodGNN.py:296-334.  vertex_attr=[b,y], edge_attr=[A_ij,c_ij], g=[n,n_A,lambda].
    The reference runs 3 GN blocks per iteration (SpMV; y=b^2 and n=sqrt(sum y); b/n) and 2 for
    the Rayleigh quotient.  Here an iteration is ONE launch of glab_power_step (SpMV + sum of
    squares reduced in-kernel; the division by n is folded into the next step) and the Rayleigh
    quotient ONE launch of glab_rayleigh.  All scalars stay on the device; nothing syncs."""

    def __init__(self, num_iter):
        super().__init__()
        self.num_iter = num_iter

    def _forward_partitioned(self, vertex_attr, pg, edge_attr, g):
        """This rank's row block of a row-partitioned operator (edgeij_pair = dist.PartitionedGraph):
        the squared norms are reduced over the ranks once per iteration."""
        io = Placement(vertex_attr, edge_attr)
        dt = float_dtype(vertex_attr, edge_attr)
        op = pg.operator(edge_attr, 1, dt)
        va = io.up(vertex_attr, dt)
        ent = op.entry()
        rt.unpack(va, [(0, 1)], outs=[op.local(ent)])
        op.publish(ent)
        res, b_out, y_out = op.power_method(self.num_iter, ent)       # res = [lambda, n, n_A] (fp64, device)
        g_dev = io.up(g).to(torch.float64)
        norm = res[1] if self.num_iter > 0 else g_dev[0]
        g_out = torch.stack([norm, res[2], res[0]]).to(g.dtype if g.dtype.is_floating_point else dt)
        op.load(ent, b_out)                        # the message column gathers the normalised iterate,
        op.acquire(ent)                            # halo rows included
        e_out = rt.with_messages(op.plan, op.vals, op.vec[ent])
        v_out = rt.pack([b_out, y_out])
        return io.down(v_out), io.down(e_out), io.down(g_out)

    def forward(self, vertex_attr, edgeij_pair, edge_attr, g, batch=None):
        from .dist import is_partitioned
        if is_partitioned(edgeij_pair):
            return self._forward_partitioned(vertex_attr, edgeij_pair, edge_attr, g)
        io = Placement(vertex_attr, edgeij_pair, edge_attr)
        dt = float_dtype(vertex_attr, edge_attr)
        n = vertex_attr.shape[0]
        plan = rt.get_plan(edgeij_pair, n)
        vals = rt.get_vals(plan, edge_attr, 0, dt)
        va = io.up(vertex_attr, dt)
        cur = rt.unpack(va, [(0, 1)])[0].view(-1)
        nxt = torch.empty_like(cur)
        sums = torch.zeros(2 * (self.num_iter + 2), dtype=torch.float64, device=io.device)
        ss_prev = None
        for it in range(self.num_iter):
            ss = sums[2 * it:2 * it + 2]
            rt.power_step(plan, vals, cur, nxt, ss_prev, ss)
            cur, nxt = nxt, cur
            ss_prev = ss
        b_out = nxt if nxt.data_ptr() != va.data_ptr() else torch.empty_like(cur)
        y_out = torch.empty_like(cur)
        ray = sums[2 * self.num_iter:2 * self.num_iter + 2]
        rt.rayleigh(plan, vals, cur, b_out, y_out, ss_prev, ray)
        g_dev = io.up(g).to(torch.float64)
        norm = torch.sqrt(ss_prev[0]) if ss_prev is not None else g_dev[0]
        g_out = torch.stack([norm, ray[0], ray[0] / ray[1]]).to(g.dtype if g.dtype.is_floating_point else dt)
        e_out = rt.with_messages(plan, vals, b_out)
        v_out = rt.pack([b_out, y_out])
        return io.down(v_out), io.down(e_out), io.down(g_out)
