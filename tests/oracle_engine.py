"""Oracle-backed compute engine for the CPU (gloo) tests of sipp_b200.sharded.  TEST ONLY: it lets the multi-rank
host logic (sharding, all-gather of partials, challenge broadcast, tail collapse) run without a GPU."""
import torch

from oracle import pyoracle as o

ONE = (1).to_bytes(32, "little") + bytes(352)


class OracleEngine:
    class _Ctx:
        def __init__(self, A, B):
            self.A, self.B = bytes(A), bytes(B)

        def __len__(self):
            return len(self.A) // 64

        def _miller_product(self, A, B):
            acc = ONE
            for i in range(len(A) // 64):
                acc = o.field_op("FQ12_MUL", acc, o.miller_loop(A[64 * i:64 * i + 64], B[128 * i:128 * i + 128]))
            return acc

        def partial_products(self, which):
            n = len(self)
            if which == 0:
                out = self._miller_product(self.A, self.B)
            else:
                h = n // 2
                out = self._miller_product(self.A[64 * h:], self.B[:128 * h]) + self._miller_product(self.A[:64 * h], self.B[128 * h:])
            return torch.frombuffer(bytearray(out), dtype=torch.uint8)

        def fold(self, x, x_inv):
            self.A, self.B = o.fold_g1(self.A, x), o.fold_g2(self.B, x_inv)

        def read(self):
            return self.A, self.B

    def create(self, A, B):
        return OracleEngine._Ctx(A, B)

    def combine(self, gathered, count, nprod):
        g = bytes(gathered.numpy().tobytes())
        outs = []
        for p in range(nprod):
            acc = ONE
            for r in range(count):
                off = 384 * (r * nprod + p)
                acc = o.field_op("FQ12_MUL", acc, g[off:off + 384])
            outs.append(o.final_exp(acc))
        return outs

    def tensor_device(self):
        return torch.device("cpu")
