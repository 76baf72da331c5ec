"""bf16-storage emulation for the CPU oracle (test infrastructure).

The B200 path keeps activations and GEMM operands in bf16 (BASELINE.json: "bf16 operands and
fp32 accumulation").  `emulate_bf16()` makes oracle/ref_port.py round its tensors to bf16 at
exactly the points where the B200 path stores them (conv outputs, BN+activation outputs that feed
a GEMM, block outputs, packed weights), keeping all arithmetic in fp32.  This separates the two
sources of deviation from the fp32 reference: the operand format (shared with the emulation) and
implementation error (what remains between the kernels and the emulation)."""
import contextlib

import torch

from oracle import ref_port as O


class _RoundBF16(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t):
        return t.to(torch.bfloat16).to(t.dtype)

    @staticmethod
    def backward(ctx, g):
        return g


@contextlib.contextmanager
def emulate_bf16(trace=None, fold_eval=True):
    """fold_eval: eval-mode BatchNorm layers are epilogues of their convolutions (the B200 inference path under
    torch.no_grad(); oracle/ref_port.py Q_FOLD_EVAL) -- the pre-BatchNorm tensors are not rounding points."""
    old = (O.Q, O.TRACE, O.Q_FOLD_EVAL)
    O.Q, O.TRACE, O.Q_FOLD_EVAL = _RoundBF16.apply, trace, fold_eval
    try:
        yield
    finally:
        O.Q, O.TRACE, O.Q_FOLD_EVAL = old


@contextlib.contextmanager
def traced(trace):
    old = O.TRACE
    O.TRACE = trace
    try:
        yield
    finally:
        O.TRACE = old
