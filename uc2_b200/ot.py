"""Wasserstein (optimal transport) distance -- mirror of model/ot.py with the IPOT loop, the cosine
cost and the trace fused into one CUDA launch per call (uc2_ot_ipot_fwd / _bwd)."""
import torch

from ._lib import call, stream

BF16, F32 = torch.bfloat16, torch.float32


class _OtFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, seq, scatter, txt_pad, img_pad, tl, beta, iteration, k):
        seq = seq.contiguous()
        B, S, H = seq.shape
        scatter = scatter.to(torch.long).contiguous()
        tp = txt_pad.to(torch.uint8).contiguous()
        ip = img_pad.to(torch.uint8).contiguous()
        M, N = tp.size(1), ip.size(1)
        dist = torch.empty(B, dtype=F32, device=seq.device)
        Cs = torch.empty((B, N, M), dtype=F32, device=seq.device)
        Ts = torch.empty((B, N, M), dtype=F32, device=seq.device)
        call("uc2_ot_ipot_fwd", seq.data_ptr(), scatter.data_ptr(), tp.data_ptr(), ip.data_ptr(), B, S, M, N, tl,
             float(beta), int(iteration), int(k), dist.data_ptr(), Cs.data_ptr(), Ts.data_ptr(), stream())
        ctx.save_for_backward(seq, scatter, tp, ip, Cs, Ts)
        ctx.tl = tl
        return dist

    @staticmethod
    def backward(ctx, ddist):
        seq, scatter, tp, ip, Cs, Ts = ctx.saved_tensors
        B, S, H = seq.shape
        dseq = torch.zeros_like(seq)
        ddist = ddist.to(F32).contiguous()
        call("uc2_ot_ipot_bwd", seq.data_ptr(), scatter.data_ptr(), tp.data_ptr(), ip.data_ptr(), B, S, tp.size(1),
             ip.size(1), ctx.tl, Cs.data_ptr(), Ts.data_ptr(), ddist.data_ptr(), dseq.data_ptr(), stream())
        return dseq, None, None, None, None, None, None, None


def optimal_transport_dist_packed(sequence_output, ot_inputs, tl, il, beta=0.5, iteration=50, k=1):
    """forward_itm's OT branch (model/model.py:701-720): un-pack through ot_scatter, then
    optimal_transport_dist(txt_emb, img_emb, txt_pad, img_pad).  Returns [B]."""
    if ot_inputs["img_pad"].size(1) > il:
        raise ValueError("img_pad is wider than the image block")
    return _OtFn.apply(sequence_output, ot_inputs["ot_scatter"], ot_inputs["txt_pad"], ot_inputs["img_pad"],
                       int(tl), beta, iteration, k)


def optimal_transport_dist(txt_emb, img_emb, txt_pad, img_pad, beta=0.5, iteration=50, k=1):
    """Same signature as model/ot.py:66: [B,M,D], [B,N,D], [B,M], [B,N] -> [B]."""
    B, M, D = txt_emb.shape
    N = img_emb.size(1)
    seq = torch.cat([txt_emb, img_emb], 1).to(BF16)
    scatter = torch.arange(M + N, device=seq.device).unsqueeze(0).expand(B, -1)
    return _OtFn.apply(seq, scatter, txt_pad, img_pad, M, beta, iteration, k)
