"""The algebraic rewrites that the fused decoder kernels rely on, checked in fp64 against the plain
cross-attention of the oracle (`DecAttention`, oracle/sam_ref.py) -- no GPU involved.

* token -> image (t2i_fused.cu): the k / v projections are folded into the query / output side,
    scores = (Wk_h^T q_h / 4) . (x + pe)      (the k bias is constant along the softmax axis)
    out_h  = Wv_h (sum_j p_j x_j) + bv_h      (the probabilities sum to one)
* image -> token (i2t_fused.cu): scores = (x + pe) . Mq + c, output = P V' with
    Mq[(h,t)] = Wq_h^T k_tok[t,h] / 4,  c[(h,t)] = bq_h . k_tok[t,h] / 4,  V'[(h,t)] = Wo_h v_tok[t,h].
"""
import torch

from oracle import sam_ref


def _attn(seed):
    torch.manual_seed(seed)
    a = sam_ref.DecAttention(256, 8, downsample_rate=2).double()
    for p in a.parameters():
        torch.nn.init.normal_(p, std=0.2)
    return a


def test_token_to_image_folding_is_exact():
    a = _attn(0)
    T, N = 7, 300
    q_in, x, pe = torch.randn(1, T, 256).double(), torch.randn(1, N, 256).double(), torch.randn(1, N, 256).double()
    ref = a(q=q_in, k=x + pe, v=x)                                           # includes out_proj
    q = a.q_proj(q_in)[0]                                                    # [T, 128]
    Wk, Wv, bv = a.k_proj.weight, a.v_proj.weight, a.v_proj.bias             # [128, 256]
    heads = []
    for h in range(8):
        sl = slice(16 * h, 16 * h + 16)
        qp = 0.25 * q[:, sl] @ Wk[sl]                                        # Q'[(h,t), :] = Wk_h^T q_h / 4      [T, 256]
        p = torch.softmax(qp @ (x[0] + pe[0]).T, dim=-1)                     # k bias dropped: softmax invariant
        u = p @ x[0]                                                         # U[(h,t), :] = sum_j p_j x_j         [T, 256]
        heads.append(u @ Wv[sl].T + bv[sl])                                  # value projection after the attention
    out = a.out_proj(torch.cat(heads, dim=-1))[None]
    assert torch.allclose(out, ref, rtol=1e-10, atol=1e-10)


def test_image_to_token_folding_is_exact():
    a = _attn(1)
    T, N = 7, 300
    tok_k, tok_v = torch.randn(1, T, 256).double(), torch.randn(1, T, 256).double()
    x, pe = torch.randn(1, N, 256).double(), torch.randn(1, N, 256).double()
    ref = a(q=x + pe, k=tok_k, v=tok_v)                                      # attention output incl. out_proj [1, N, 256]
    k_tok, v_tok = a.k_proj(tok_k)[0], a.v_proj(tok_v)[0]                    # [T, 128]
    Wq, bq, Wo, bo = a.q_proj.weight, a.q_proj.bias, a.out_proj.weight, a.out_proj.bias
    Mq, c, Vp = [], [], []
    for h in range(8):
        sl = slice(16 * h, 16 * h + 16)
        Mq.append(0.25 * k_tok[:, sl] @ Wq[sl])                              # [T, 256]
        c.append(0.25 * k_tok[:, sl] @ bq[sl])                               # [T]
        Vp.append(v_tok[:, sl] @ Wo[:, sl].T)                                # [T, 256]
    s = torch.stack([(x[0] + pe[0]) @ Mq[h].T + c[h] for h in range(8)])     # [8, N, T]
    p = torch.softmax(s, dim=-1)
    out = sum(p[h] @ Vp[h] for h in range(8)) + bo                           # [N, 256]
    assert torch.allclose(out[None], ref, rtol=1e-10, atol=1e-10)
    # the bias can NOT be dropped here (softmax over the tokens): doing so changes the result
    s0 = torch.stack([(x[0] + pe[0]) @ Mq[h].T for h in range(8)])
    assert not torch.allclose(sum(torch.softmax(s0, -1)[h] @ Vp[h] for h in range(8)) + bo, out, atol=1e-6)
