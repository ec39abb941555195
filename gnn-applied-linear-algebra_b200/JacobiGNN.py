"""Drop-in for pytorch/JacobiGNN.py: weighted Jacobi x <- x + w (b - A x) / A_ii."""
import torch

from . import _runtime as rt
from ._io import Placement, float_dtype


class JacobiGNN(torch.nn.Module):
    """JacobiGNN.py:125-148.  vertex_attr=[A_ii, b, x], edge_attr=[A_ij, c_ij] (edges INCLUDE
    the diagonal), g=[w].  Each sweep is one launch of glab_jacobi (gather + multiply +
    row sum + update fused); sweeps ping-pong between two x buffers, and forward() runs all n_iters
    sweeps in one launch of the multi-sweep kernel (glab_jacobi_sweeps_*).
    Extension: vertex_attr = [A_ii | b (k cols) | x (k cols)] smooths k right-hand sides."""

    def _setup(self, vertex_attr, edgeij_pair, edge_attr, g):
        io = Placement(vertex_attr, edgeij_pair, edge_attr)
        dt = float_dtype(vertex_attr, edge_attr)
        n, F = vertex_attr.shape
        k = (F - 1) // 2
        plan = rt.get_plan(edgeij_pair, n)
        vals = rt.get_vals(plan, edge_attr, 0, dt)
        va = io.up(vertex_attr, dt)
        diag, b, x = rt.unpack(va, [(0, 1), (1, k), (1 + k, k)])
        diag = diag.view(-1)
        w = rt.scalar(g[0] if isinstance(g, torch.Tensor) else g, io.device, dt)
        return io, dt, plan, vals, va, diag, b, x, w

    def iterate(self, vertex_attr, edgeij_pair, edge_attr, g, batch=None):
        io, dt, plan, vals, va, diag, b, x, w = self._setup(vertex_attr, edgeij_pair, edge_attr, g)
        x_new = rt.jacobi(plan, vals, diag, b, x, torch.empty_like(x), w)
        e_out = rt.with_messages(plan, vals, x)
        v_out = rt.pack([diag, b, x_new])
        return io.down(v_out), io.down(e_out), g

    def _forward_partitioned(self, n_iters, vertex_attr, pg, edge_attr, g):
        """This rank's row block of a row-partitioned operator (edgeij_pair = dist.PartitionedGraph):
        the same sweeps, halo rows pushed from inside the kernels."""
        io = Placement(vertex_attr, edge_attr)
        dt = float_dtype(vertex_attr, edge_attr)
        n, F = vertex_attr.shape
        k = (F - 1) // 2
        op = pg.operator(edge_attr, k, dt)
        va = io.up(vertex_attr, dt)
        ent = op.entry()
        diag, b, _ = rt.unpack(va, [(0, 1), (1, k), (1 + k, k)], outs=[None, None, op.local(ent)])
        w = rt.scalar(g[0] if isinstance(g, torch.Tensor) else g, io.device, dt)
        op.publish(ent)
        cur = op.jacobi(n_iters, diag.view(-1), b, w, ent)
        return io.down(op.local(cur).clone())

    def forward(self, n_iters, vertex_attr, edgeij_pair, edge_attr, g, batch=None):
        from .dist import is_partitioned
        if is_partitioned(edgeij_pair):
            return self._forward_partitioned(n_iters, vertex_attr, edgeij_pair, edge_attr, g)
        io, dt, plan, vals, va, diag, b, x, w = self._setup(vertex_attr, edgeij_pair, edge_attr, g)
        # all sweeps in one launch of the multi-sweep kernel (x is a private copy from unpack)
        x = rt.jacobi_sweeps(plan, vals, diag, b, x, torch.empty_like(x), w, n_iters)
        return io.down(x.reshape(x.shape[0], -1))
