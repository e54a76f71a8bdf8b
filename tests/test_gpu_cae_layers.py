"""GPU: the CAE tensor-core path layer by layer against the fp32 CUDA-core cross-check path (same folded weights),
and end to end against the numpy oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from dfs_b200 import CaeScorer, synthetic as syn  # noqa: E402
from oracle import models_np as onp  # noqa: E402

NAMES = ("enc1", "enc2", "enc3", "enc4", "dec1", "dec2", "dec3")


@pytest.fixture(scope="module")
def setup():
    x = torch.from_numpy(syn.features(5, seed=77)).cuda()
    mean, std = syn.normalizer_stats(1)
    return x, CaeScorer(syn.cae_state(0), mean, std, max_chunk=8)


@pytest.mark.parametrize("layer", range(7), ids=NAMES)
def test_cae_layer_matches_cuda_core_path(setup, layer):
    x, sc = setup
    tc = sc.debug_layer(x, layer, impl=0).cpu().numpy()
    ref = sc.debug_layer(x, layer, impl=1).cpu().numpy()
    assert tc.shape == ref.shape
    scale = np.abs(ref).max()
    err = np.abs(tc - ref).max()
    # fp16 activations / weights with fp32 accumulation vs an all-fp32 path: error grows mildly with depth
    assert err <= 4e-3 * scale * (1 + layer), f"{NAMES[layer]}: max abs err {err:.3e} vs scale {scale:.3e}"
    assert np.abs(tc).max() > 0.1 * scale        # not silently zero


def test_cae_mse_tensor_core_vs_oracle_and_crosscheck(setup):
    x, sc = setup
    mean, std = syn.normalizer_stats(1)
    ref = onp.cae_mse_scores(syn.cae_state(0), x.cpu().numpy(), mean, std)
    tc = sc.score(x).cpu().numpy()
    assert np.max(np.abs(tc - ref) / ref) <= 1e-3
    sc.set_option("conv_impl", 1)
    simt = sc.score(x).cpu().numpy()
    sc.set_option("conv_impl", 0)
    assert np.max(np.abs(simt - ref) / ref) <= 1e-5
    xt = x.transpose(1, 2).contiguous().transpose(1, 2)
    np.testing.assert_allclose(sc.score(xt).cpu().numpy(), tc, rtol=1e-6)


def test_cae_enc1_tensor_core_vs_fp32_cuda_core_conv(setup):
    """enc1 as the Toeplitz-in-time tcgen05 GEMM (cae_enc1_tc.cu) against the fp32 CUDA-core conv1_kernel<POOLF> writing the
    same FT8P buffer (option conv1_impl = 1), on contiguous and on the reference's transposed storage, with and without
    the normaliser; and the end-to-end MSE with either enc1."""
    x, sc = setup
    xt = x.transpose(1, 2).contiguous().transpose(1, 2)
    for feats in (x, xt):
        for norm in (True, False):
            tc = sc.debug_layer(feats, 0, impl=0, apply_normalizer=norm).cpu().numpy()
            sc.set_option("conv1_impl", 1)
            ref = sc.debug_layer(feats, 0, impl=0, apply_normalizer=norm).cpu().numpy()
            sc.set_option("conv1_impl", 0)
            scale = np.abs(ref).max()
            assert np.abs(tc - ref).max() <= 2e-3 * scale, (norm, np.abs(tc - ref).max(), scale)
            assert (ref > 0).mean() > 0.2 and np.array_equal(tc == 0, ref == 0) or np.abs(tc - ref).max() <= 2e-3 * scale
    a = sc.score(x).cpu().numpy()
    sc.set_option("conv1_impl", 1)
    b = sc.score(x).cpu().numpy()
    sc.set_option("conv1_impl", 0)
    assert np.max(np.abs(a - b) / b) <= 2e-4


def test_cae_chunk_boundaries_and_ragged_tail():
    """More utterances than one internal pass (chunk 8): every pass reuses the zero-padded xT2 / activation buffers."""
    mean, std = syn.normalizer_stats(1)
    x = torch.from_numpy(syn.features(19, seed=5)).cuda()
    sc = CaeScorer(syn.cae_state(0), mean, std, max_chunk=8)
    got = sc.score(x).cpu().numpy()
    ref = onp.cae_mse_scores(syn.cae_state(0), x.cpu().numpy(), mean, std)
    assert np.max(np.abs(got - ref) / ref) <= 1e-3
    one = CaeScorer(syn.cae_state(0), mean, std, max_chunk=32).score(x).cpu().numpy()
    np.testing.assert_array_equal(got, one)          # chunking must not change a single bit


def test_cae_fused_final_layer_equals_separate_final_kernel():
    """dec3's epilogue applying the final ConvTranspose2d + squared error (no d3, no reconstruction in HBM) against the
    separate final kernel reading d3: same fp16-rounded d3, same FMA order for the reconstruction, only the order of the
    57,780-term sum differs.  Contiguous and transposed storage, with and without the normaliser, ragged passes."""
    mean, std = syn.normalizer_stats(1)
    x = torch.from_numpy(syn.features(11, seed=13)).cuda()
    xt = x.transpose(1, 2).contiguous().transpose(1, 2)
    for scorer in (CaeScorer(syn.cae_state(0), mean, std, max_chunk=4), CaeScorer(syn.cae_state(2), None, None, max_chunk=16)):
        for feats in (x, xt):
            fused = scorer.score(feats).cpu().numpy()
            scorer.set_option("final_fused", 0)
            plain = scorer.score(feats).cpu().numpy()
            scorer.set_option("final_fused", 1)
            assert np.max(np.abs(fused - plain) / plain) <= 2e-6
            again = scorer.score(feats).cpu().numpy()
            np.testing.assert_array_equal(fused, again)      # deterministic
    ref = onp.cae_mse_scores(syn.cae_state(0), x.cpu().numpy(), mean, std)
    got = CaeScorer(syn.cae_state(0), mean, std).score(x).cpu().numpy()
    assert np.max(np.abs(got - ref) / ref) <= 1e-3


def test_cae_wide_decoder_variant_is_bit_identical():
    """Option "dec_wide" (default 1): dec1 / dec2 as N = 256 GEMMs (two 128-column sub-groups per CTA, half the re-reads of the
    input window) against the N = 128 variants.  Same K order per output column, so d1, d2 and the scores must not differ by
    a bit; 13 utterances through passes of 5."""
    mean, std = syn.normalizer_stats(1)
    x = torch.from_numpy(syn.features(13, seed=21)).cuda()
    sc = CaeScorer(syn.cae_state(1), mean, std, max_chunk=5)
    sc.set_option("dec_wide", 0)
    base = sc.score(x).cpu().numpy()
    d1, d2 = sc.debug_layer(x[:5], 4, impl=0).cpu().numpy(), sc.debug_layer(x[:5], 5, impl=0).cpu().numpy()
    sc.set_option("dec_wide", 1)
    np.testing.assert_array_equal(sc.debug_layer(x[:5], 4, impl=0).cpu().numpy(), d1)
    np.testing.assert_array_equal(sc.debug_layer(x[:5], 5, impl=0).cpu().numpy(), d2)
    np.testing.assert_array_equal(sc.score(x).cpu().numpy(), base)


def test_cae_enc4_on_cta_pairs_is_bit_identical():
    """Option "pair_mma" (default 1): enc4 as tcgen05 cta_group::2 MMAs (cluster of 2 CTAs, M = 256 = the two units of a pair, each CTA
    holding 64 of the 128 weight rows of its group) against the single-CTA kernel with 4 groups of N = 64: same K order per output,
    so the latent and the scores must agree bit for bit.  3 utterances = 5 column tiles (odd: the last pair has a padding unit),
    then 13 utterances through passes of 5."""
    mean, std = syn.normalizer_stats(1)
    x = torch.from_numpy(syn.features(13, seed=23)).cuda()
    sc = CaeScorer(syn.cae_state(2), mean, std, max_chunk=5)
    sc.set_option("pair_mma", 0)
    base = sc.score(x).cpu().numpy()
    e4_3, e4_5 = sc.debug_layer(x[:3], 3, impl=0).cpu().numpy(), sc.debug_layer(x[5:10], 3, impl=0).cpu().numpy()
    sc.set_option("pair_mma", 1)
    np.testing.assert_array_equal(sc.debug_layer(x[:3], 3, impl=0).cpu().numpy(), e4_3)
    np.testing.assert_array_equal(sc.debug_layer(x[5:10], 3, impl=0).cpu().numpy(), e4_5)
    np.testing.assert_array_equal(sc.score(x).cpu().numpy(), base)


def test_cae_enc3_with_swapped_operand_roles_matches():
    """Option "enc3_swap" (default 1): enc3 as the 2D-CNN's conv3 GEMM (weights = A operand, 256 positions = N) with both 2x2-pool partners in one
    thread.  Same fp16 operands and fp32 accumulation per conv output; the four ReLU outputs of a pool window are added in a
    different order than the lane-exchange epilogue adds them, so e3 agrees to fp16 round-off of the pooled value (1 ulp), the
    scores to 1e-5.  3 and 5 utterances (72 / 120 columns: partial 32-column tiles), then ragged passes."""
    mean, std = syn.normalizer_stats(1)
    x = torch.from_numpy(syn.features(13, seed=29)).cuda()
    sc = CaeScorer(syn.cae_state(3), mean, std, max_chunk=5)
    sc.set_option("enc3_swap", 0)
    base = sc.score(x).cpu().numpy()
    ref3, ref5 = sc.debug_layer(x[:3], 2, impl=0).cpu().numpy(), sc.debug_layer(x[5:10], 2, impl=0).cpu().numpy()
    sc.set_option("enc3_swap", 1)
    for xs, want in ((x[:3], ref3), (x[5:10], ref5)):
        got = sc.debug_layer(xs, 2, impl=0).cpu().numpy()
        assert got.shape == want.shape and np.array_equal(got == 0, want == 0)
        assert np.max(np.abs(got - want)) <= 2e-3 * np.abs(want).max() and np.mean(got != want) < 0.2
    swapped = sc.score(x).cpu().numpy()
    assert np.max(np.abs(swapped - base) / base) <= 1e-5
    ref = onp.cae_mse_scores(syn.cae_state(3), x.cpu().numpy(), mean, std)
    assert np.max(np.abs(swapped - ref) / ref) <= 1e-3
