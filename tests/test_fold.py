"""Host logic: the folded relation weights reproduce the unfolded reference chain."""
import torch
import torch.nn.functional as F

from faster_orefsdet_b200 import fold, synth
from oracle import head_oracle as O
from tests.util import assert_close, head_state_dict


def test_folded_relation_head_equals_unfolded():
    sd = head_state_dict()
    x = synth.tensor((40, 128, 8, 8), 5, -2.0, 2.0)
    sup = synth.tensor((2, 6, 128, 8, 8), 6, -1.0, 1.0)
    w_fold, w_out, b_out = fold.fold_relation_weights(sd)
    bias = fold.fold_class_bias(sd, sup.mean(1))
    for c in range(2):
        logits, deltas = O.relation_head(x, sup[c], sd)
        xk = x.permute(0, 2, 3, 1).reshape(40, 8192)            # [roi][bin][channel]
        f = F.relu(xk.double() @ w_fold.double().t() + bias[c].double())
        out = (f @ w_out.double().t() + b_out.double()).float()
        assert_close(out[:, :2], logits, rtol=1e-4, atol=1e-5, what="logits")
        assert_close(out[:, 2:], deltas, rtol=1e-4, atol=1e-5, what="deltas")
