"""GatherLayer (detectron2/modeling/backbone/clipcap/gather.py:5-20): differentiable all-gather whose backward
keeps only the local slice of the gradient (no reduction).  Pure `torch.distributed`, so it runs on NCCL
(GPU) and gloo (CPU tests) alike.  The fused path in `caption_consistency.py` does not use it — it gathers
one packed, already-normalised buffer instead — but the symbol is part of the reference's surface."""
import torch
import torch.distributed as dist


class GatherLayer(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input):
        ctx.save_for_backward(input)
        output = [torch.zeros_like(input) for _ in range(dist.get_world_size())]
        dist.all_gather(output, input.contiguous())
        return tuple(output)

    @staticmethod
    def backward(ctx, *grads):
        (input,) = ctx.saved_tensors
        grad_out = torch.zeros_like(input)
        grad_out[:] = grads[dist.get_rank()]
        return grad_out
