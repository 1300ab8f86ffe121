"""Pivot translator (unpaired_image_captioning_b200/pivot.py): the shape-static masked bi-LSTM equals torch's packed
nn.LSTM (outputs, final states, gradients), and the whole step is a function of the padded batch only."""
import pytest
import torch
import torch.nn as nn

from unpaired_image_captioning_b200 import pivot


def test_masked_bilstm_equals_packed_lstm():
    torch.manual_seed(3)
    S, B, D, Hh = 9, 5, 12, 8
    enc = pivot.MaskedBiLSTM(D, Hh, num_layers=2, dropout=0.0)
    x = torch.randn(S, B, D, requires_grad=True)
    lengths = torch.tensor([9, 7, 7, 4, 1])
    mem, (h, c) = enc(x, lengths)
    (mem.sum() + h.sum() * 0.5 + c.sum() * 0.25).backward()
    g_x, g_w = x.grad.clone(), enc.rnn.weight_hh_l1_reverse.grad.clone()
    x.grad = None
    enc.zero_grad()
    packed = nn.utils.rnn.pack_padded_sequence(x, lengths)
    out, (h2, c2) = enc.rnn(packed)
    mem2 = nn.utils.rnn.pad_packed_sequence(out, total_length=S)[0]
    (mem2.sum() + h2.sum() * 0.5 + c2.sum() * 0.25).backward()
    torch.testing.assert_close(mem, mem2, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(h, h2, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(c, c2, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(g_x, x.grad, rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(g_w, enc.rnn.weight_hh_l1_reverse.grad, rtol=1e-4, atol=1e-6)


def test_translator_loss_ignores_padding_beyond_the_lengths():
    torch.manual_seed(4)
    gen = torch.Generator().manual_seed(4)
    m = pivot.PivotNMT(src_vocab=50, tgt_vocab=40, dim=16, layers=2, dropout=0.0).eval()
    src, n = pivot.sentences(4, 50, gen, lo=2, hi=6)
    tgt, _ = pivot.sentences(4, 40, gen, lo=2, hi=6, bos=pivot.BOS)
    nll, cnt = m(src, n, tgt)
    src_pad = torch.cat([src, torch.zeros(3, 4, dtype=torch.int64)], 0)            # a longer padded batch: same sentences
    tgt_pad = torch.cat([tgt, torch.zeros(2, 4, dtype=torch.int64)], 0)
    nll2, cnt2 = m(src_pad, n, tgt_pad)
    assert int(cnt) == int(cnt2) == int((tgt[1:] != pivot.PAD).sum())
    torch.testing.assert_close(nll, nll2, rtol=1e-5, atol=1e-5)


def test_translator_equals_the_plain_per_token_formulation():
    """The kernel-count restructuring of PivotNMT.forward (word half of the first layer's input projection time-batched, one
    GEMM over [feed | h0] per step, additive score mask) computes what the reference's step-by-step formulation does:
    input-feed stacked LSTMCells on cat([emb_t, feed]) + "general" global attention (models/NMT_Models.py:209-262)."""
    import torch.nn.functional as F
    torch.manual_seed(0)
    gen = torch.Generator().manual_seed(1)
    m = pivot.PivotNMT(src_vocab=50, tgt_vocab=40, dim=16, layers=2, dropout=0.0).eval()
    src, n = pivot.sentences(5, 50, gen, lo=2, hi=7)
    tgt, _ = pivot.sentences(5, 40, gen, lo=2, hi=7, bos=pivot.BOS)
    nll, cnt = m(src, n, tgt)
    nll.backward()
    g = {k: p.grad.clone() for k, p in m.named_parameters()}
    m.zero_grad()

    emb = F.relu(m.src_mlp(m.src_lut(src)))
    memory, (h, c) = m.encoder(emb, n)
    memory = memory.transpose(0, 1)
    fix = lambda s: torch.cat([s[0::2], s[1::2]], 2)
    h, c = list(fix(h)), list(fix(c))
    mask = torch.arange(memory.size(1))[None, :] >= n[:, None]
    feed, keys, outs = memory.new_zeros(src.size(1), m.dim), m.attn_in(memory), []
    te = m.tgt_lut(tgt[:-1])
    for t in range(te.size(0)):
        x = torch.cat([te[t], feed], 1)
        for i, cell in enumerate(m.cells):
            h[i], c[i] = cell(x, (h[i], c[i]))
            x = h[i]
        score = torch.bmm(keys, x.unsqueeze(2)).squeeze(2).masked_fill(mask, float("-inf"))
        ctx = torch.bmm(F.softmax(score, 1).unsqueeze(1), memory).squeeze(1)
        feed = torch.tanh(m.attn_out(torch.cat([ctx, x], 1)))
        outs.append(feed)
    logp = F.log_softmax(m.generator(torch.stack(outs)), -1)
    ref = F.nll_loss(logp.view(-1, logp.size(-1)), tgt[1:].reshape(-1), ignore_index=pivot.PAD, reduction="sum")
    ref.backward()
    torch.testing.assert_close(nll, ref, rtol=1e-5, atol=1e-5)
    for k, p in m.named_parameters():
        torch.testing.assert_close(g[k], p.grad, rtol=1e-4, atol=1e-5, msg=k)


def test_input_feed_decoder_backward_with_dropout_gradcheck():
    """The hand-written backward of the decoder loop (pivot._InputFeedDecoderFn) with ACTIVE dropout against numerical
    derivatives (float64; the generator is re-seeded inside the function, so the masks are a fixed part of it)."""
    torch.manual_seed(11)
    T, B, d, S, L = 3, 2, 4, 3, 2
    f64 = dict(dtype=torch.float64, requires_grad=True)
    eg = torch.randn(T, B, 4 * d, **f64)
    w_fh = (torch.randn(4 * d, 2 * d, dtype=torch.float64) * 0.5).requires_grad_()
    keys, memory = torch.randn(B, S, d, **f64), torch.randn(B, S, d, **f64)
    neg = torch.zeros(B, S, dtype=torch.float64)
    neg[1, 2] = float("-inf")
    w_out = (torch.randn(d, 2 * d, dtype=torch.float64) * 0.5).requires_grad_()
    h0s, c0s = torch.randn(L, B, d, **f64), torch.randn(L, B, d, **f64)
    w_cat = (torch.randn(4 * d, 2 * d, dtype=torch.float64) * 0.5).requires_grad_()
    b1 = torch.randn(4 * d, **f64)

    def fn(eg, w_fh, keys, memory, w_out, h0s, c0s, w_cat, b1):
        torch.manual_seed(99)
        return pivot._InputFeedDecoderFn.apply(0.3, True, eg, w_fh, keys, memory, neg, w_out, h0s, c0s, w_cat, b1)

    assert torch.autograd.gradcheck(fn, (eg, w_fh, keys, memory, w_out, h0s, c0s, w_cat, b1), eps=1e-6, atol=1e-5, rtol=1e-4)


@pytest.mark.gpu
def test_graphed_translator_step_equals_eager():
    """The CUDA-graph replay of the training step follows the same loss trajectory as the step issued eagerly (dropout 0)."""
    gen = torch.Generator().manual_seed(5)
    src, n = pivot.sentences(16, 300, gen, lo=3, hi=9)
    tgt, _ = pivot.sentences(16, 200, gen, lo=3, hi=9, bos=pivot.BOS)
    losses = {}
    for graph in (False, True):
        torch.manual_seed(6)
        m = pivot.PivotNMT(src_vocab=300, tgt_vocab=200, dim=64, layers=2, dropout=0.0).cuda().train()
        step = pivot.PivotTrainStep(m, src.size(0), tgt.size(0), 16, graph=graph, batch=(src.cuda(), n.cuda(), tgt.cuda()))
        if not graph:
            for _ in range(3):          # the graphed variant warms up with three eager steps before it captures
                step.step()
        losses[graph] = [float(step.step()) for _ in range(4)]
    torch.testing.assert_close(torch.tensor(losses[True]), torch.tensor(losses[False]), rtol=2e-3, atol=2e-3)
    assert losses[True][-1] < losses[True][0]
