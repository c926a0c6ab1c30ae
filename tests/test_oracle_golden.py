"""CPU: the oracle against the committed golden vectors (produced by the unmodified reference, see
oracle/make_golden.py) and against the reference's own self-consistency invariants."""
import itertools

import numpy as np
import pytest
import torch

from helpers import CONFIGS, golden_case, make_audio, model_config
from l3ac_b200.spec import init_state_dicts
from oracle import l3ac_oracle as O


@pytest.mark.parametrize("name", ["0k75bps", "1k5bps", "3kbps", "rotary"])
def test_oracle_matches_reference_golden(name):
    mc, weights, audio, g = golden_case(name)
    orc = O.Oracle(mc.as_dict(), weights)
    taps = {}
    q, idx = orc.encode_audio(audio, taps)
    assert np.array_equal(idx["indices"].numpy(), g["indices"])
    assert np.array_equal(idx["level_indices"].numpy().astype(np.int8), g["level_indices"])
    assert np.array_equal(taps["z"].numpy(), g["z"])
    wav = orc.decode_audio(indices=idx["indices"])
    stride = int(g["wav_stride"])
    assert np.array_equal(wav.numpy()[:, ::stride], g["wav"])
    assert torch.equal(wav, orc.decode_audio(q))          # decode(q_feature) == decode(indices=), invariant (i)


def test_oracle_1kbps_encode_golden():
    mc, weights, audio, g = golden_case("1kbps")
    orc = O.Oracle(mc.as_dict(), weights)
    _, idx = orc.encode_audio(audio)
    assert np.array_equal(idx["indices"].numpy(), g["indices"])


@pytest.mark.parametrize("levels", [(7, 7, 7, 7, 7, 7), (9, 9, 9, 7, 7, 7)])
def test_fsq_codebook_round_trip_exhaustive(levels):
    """index -> codes -> index over the whole implicit codebook (117,649 / 250,047 entries)."""
    n = int(np.prod(levels))
    idx = torch.arange(n, dtype=torch.int32)
    codes = O.fsq_indices_to_codes(idx, levels)                     # values in {-1, ..., 1}
    z = torch.atanh(codes.double().clamp(-1 + 1e-12, 1 - 1e-12)).float()   # pre-activation that lands on the code
    q_z, idx2, lvl = O.fsq_quantize(z, levels)
    assert torch.equal(idx2, idx)
    assert torch.equal(q_z, codes)
    digits = torch.tensor(list(itertools.islice(itertools.product(*[range(l) for l in reversed(levels)]), 50)))
    assert torch.equal(lvl[:50].long(), digits.flip(-1))            # dimension 0 is least significant


def test_fsq_round_half_even():
    levels = (7, 7, 7, 7, 7, 7)
    # act * 6 == 2.5 exactly -> level 2 (half to even); 3.5 -> 4
    act = torch.tensor([2.5 / 6, 3.5 / 6])
    assert torch.equal((act * 6).round(), torch.tensor([2., 4.]))


def test_batch_independence_and_zero_extension():
    mc = model_config("1k5bps")
    weights = init_state_dicts(mc, seed=3, jitter=True)
    orc = O.Oracle(mc.as_dict(), weights)
    audio = make_audio(2, 2.0, seed=5)
    _, both = orc.encode_audio(audio)
    _, one = orc.encode_audio(audio[1:])
    assert torch.equal(both["indices"][1:], one["indices"])         # invariant (iii)
    _, ext = orc.encode_audio(torch.nn.functional.pad(audio[:1], (0, 1800)))
    n = both["indices"].shape[1]
    agree = (ext["indices"][:, :n - 2] == both["indices"][:1, :n - 2]).float().mean().item()
    assert agree > 0.98                                             # invariant (v): causal, right-extension is benign


@pytest.mark.parametrize("name", CONFIGS)
def test_token_rates(name):
    mc = model_config(name)
    hop = {"0k75bps": 360, "1kbps": 270, "1k5bps": 180, "3kbps": 96}[name]
    assert mc.hop_length == hop == O.hop_length(mc.as_dict())
    padded, n = O.preprocess(mc.as_dict(), torch.zeros(1, 160000))
    assert n == 160000 and padded.shape[-1] % hop == 0 and padded.shape[-1] - 160000 < hop
