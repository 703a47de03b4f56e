"""CPU: host-side mirror of the reference interface -- state_dict layout, constructor checks, tokenisation, sharding and
the world_size-2 gather (gloo)."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

import weights as W
from worddiffusion_b200 import diffusion as D
from worddiffusion_b200.unet import UNetModel, default_args
from worddiffusion_b200.unetPhosc import UNetModelPhosc
from worddiffusion_b200.unetPhosc2 import UNetModelPhosc as UNetModelPhosc2

KW = dict(image_size=(64, 256), in_channels=4, model_channels=320, out_channels=4, num_res_blocks=1,
          attention_resolutions=(1, 1), channel_mult=(1, 1), num_heads=4, num_classes=339, context_dim=320,
          vocab_size=53, max_seq_len=10)


@pytest.mark.parametrize("cls,variant", [(UNetModel, "unet"), (UNetModelPhosc, "unetPhosc"), (UNetModelPhosc2, "unetPhosc")])
def test_state_dict_layout_matches_reference(cls, variant):
    """Same keys, same order, same shapes as the reference modules (spec dumped from them by oracle/make_golden.py)."""
    m = cls(args=default_args("cpu"), **KW)
    got = [(k, tuple(v.shape)) for k, v in m.state_dict().items()]
    assert got == W.load_spec(variant)
    sd = W.make_state_dict(W.load_spec(variant))
    m.load_state_dict(sd, strict=True)  # regenerateFromtrain2.py:1214 loads strict=True


def test_fresh_model_has_reference_zero_init():
    m = UNetModel(args=default_args("cpu"), **KW)
    sd = m.state_dict()
    for k in ("out.2.weight", "out.2.bias", "input_blocks.1.0.out_layers.3.weight", "input_blocks.1.1.proj_out.weight"):
        assert float(sd[k].abs().max()) == 0.0, k  # zero_module, unet.py:152-158
    assert float(sd["input_blocks.1.0.in_layers.2.weight"].abs().max()) > 0


def test_flag_variant_layouts_match_reference():
    """args.ocrTraining / args.charImages change the reference's key set (unet.py:1468,1217-1223): same keys, order and shapes as
    the reference modules built with those flags (tests/golden/state_dict_spec_unet_variants.json, oracle/make_golden_variants.py)."""
    import json
    import os
    with open(os.path.join(os.path.dirname(__file__), "golden", "state_dict_spec_unet_variants.json")) as f:
        specs = json.load(f)
    for name, flags in (("ocr", dict(attentionMaps=1, ocrTraining=1)), ("charimg", dict(charImages=1))):
        m = UNetModel(args=default_args("cpu", **flags), **KW)
        got = [[k, list(v.shape)] for k, v in m.state_dict().items()]
        assert got == specs[name], name
    for flags in (dict(charLevelEmb=1), dict(wrdChrWrStyl=1), dict(interpolation=True)):
        m = UNetModel(args=default_args("cpu", **flags), **KW)
        assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == W.load_spec("unet")


def test_unsupported_configurations_fail_loudly():
    with pytest.raises(NotImplementedError):
        UNetModel(args=default_args("cpu"), **dict(KW, use_scale_shift_norm=True))
    with pytest.raises(NotImplementedError):
        UNetModelPhosc(args=default_args("cpu", interpolation=True), **KW)


def test_label_padding_follows_reference():
    # train.py:42-52: index + 1, padded with 52 to MAX_CHARS
    assert D.label_padding("Ab") == [1, 28] + [52] * 8
    assert D.label_padding("z")[0] == 52  # the reference's 'z'/PAD collision is preserved
    with pytest.raises(ValueError):
        D.label_padding("abcdefghijk")


def test_shard_bounds_partition():
    for n in (1, 7, 256, 1024, 1025):
        for w in (1, 2, 4, 8):
            spans = [D.shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_schedule_and_coefficients():
    d = D.Diffusion(noise_steps=1000, device="cpu")
    assert d.beta.shape == (1000,)
    i = 500
    a, ah, b = d.alpha[i], d.alpha_hat[i], d.beta[i]
    c = d._ddpm_coef[i]
    assert c[0] == float(1 / torch.sqrt(a)) and c[1] == float((1 - a) / torch.sqrt(1 - ah)) and c[2] == float(torch.sqrt(b))
    assert d.ddim_timesteps(50)[0] == 980 and d.ddim_timesteps(50)[-1] == 0
    x = torch.randn(3, 4, 8, 32)
    # noise_images is a kernel of the library (wd_noise_images): a CPU tensor fails loudly, there is no CPU fallback
    # (the arithmetic itself is checked on the GPU in tests/test_gpu_train.py)
    with pytest.raises(D.WdError):
        d.noise_images(x, torch.tensor([1, 10, 999]))
    t = d.sample_timesteps(64)
    assert int(t.min()) >= 1 and int(t.max()) <= 999


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gather_worker(rank, world, port, n_total, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = torch.arange(n_total * 4 * 8 * 32, dtype=torch.float32).reshape(n_total, 4, 8, 32)
    lo, hi = D.shard_bounds(n_total, world, rank)
    out = D.all_gather_latents(full[lo:hi].clone(), n_total, world)
    q.put((rank, bool(torch.equal(out, full))))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [8, 7])
def test_all_gather_latents_world2_gloo(n_total):
    """The only collective on the sampling path: every rank ends with all N latents in global order (ragged split too)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


# ---------------------------------------------------------------------------------------------------------------
# training, host side (train.py --ddp): batch sharding and the gradient exchange, world_size 2 over gloo
# ---------------------------------------------------------------------------------------------------------------
def test_train_shard_batch_partitions_the_global_batch():
    from worddiffusion_b200.training import shard_batch
    for n, world in [(224, 8), (224, 1), (10, 4), (7, 2), (3, 8)]:
        spans = [shard_batch(n, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    assert shard_batch(224, 3, 8) == (84, 112)  # BASELINE config 4: 28 latents per GPU


def _allreduce_worker(rank, world, port, q):
    import torch.distributed as dist
    from worddiffusion_b200.training import allreduce_sum_
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    flat = torch.full((1000,), float(rank + 1))
    ws = allreduce_sum_(flat)
    q.put((rank, ws, bool(torch.equal(flat, torch.full((1000,), 3.0)))))
    dist.destroy_process_group()


def test_train_gradient_allreduce_world2_gloo():
    """Gradient exchange of the training step: in-place SUM of the flat gradient buffer; the optimizer kernel divides by the
    returned world size (DDP averaging)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_allreduce_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, 2, True), (1, 2, True)]


def test_train_allreduce_is_identity_without_process_group():
    from worddiffusion_b200.training import allreduce_sum_
    flat = torch.ones(10)
    assert allreduce_sum_(flat) == 1 and torch.equal(flat, torch.ones(10))


def test_training_module_refuses_cpu():
    """No CPU fallback on the training path either."""
    from worddiffusion_b200 import _lib
    from worddiffusion_b200.unet import UNetModel, default_args
    m = UNetModel(image_size=(64, 256), in_channels=4, model_channels=320, out_channels=4, num_res_blocks=1,
                  attention_resolutions=(1, 1), channel_mult=(1, 1), num_heads=4, num_classes=339, context_dim=320,
                  vocab_size=53, max_seq_len=10, args=default_args("cpu"))
    m.train()
    x = torch.zeros(1, 4, 8, 32)
    with pytest.raises(_lib.WdError):
        m(x, None, timesteps=torch.tensor([5]), context=torch.ones(1, 10, dtype=torch.long), y=torch.tensor([1]))


def test_reduced_call_schedule_matches_oracle():
    """The evaluation schedule of the reduced-call sampler (regenerateFromtrain2.py:536) is host logic: same predicate as the oracle
    for every step of the reference's T = 1000 and T = 600 schedules; about a fifth of the steps evaluate the UNet."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    from diffusion_oracle import DiffusionOracle
    from worddiffusion_b200.diffusion import Diffusion
    for T in (1000, 600, 12):
        ours = [i for i in range(1, T) if Diffusion.reduced_call_predicate(i, T)]
        ref = [i for i in range(1, T) if DiffusionOracle.reduced_call_predicate(i, T)]
        assert ours == ref and (T - 1) in ours
    assert 0.19 < len([i for i in range(1, 1000) if Diffusion.reduced_call_predicate(i, 1000)]) / 999 < 0.21


def test_phosc_tokenizer_refuses_cpu_and_validates_words():
    from worddiffusion_b200._lib import WdError
    from worddiffusion_b200.phosc import PHOSC_LEN, phosc_labels
    assert PHOSC_LEN == 165 + 604
    with pytest.raises(WdError):
        phosc_labels(["word"], "cpu")


def test_precision_switch_host_logic():
    """`precision` is an attribute (the reference constructor has no such argument): default bf16; both engines refuse a CPU device;
    the wd_config built for either engine mirrors the constructor arguments (unet.py:1126-1156)."""
    from worddiffusion_b200 import _lib
    from worddiffusion_b200.engine import F32Engine, make_config
    m = UNetModel(args=default_args("cpu"), **KW)
    assert m.precision == "bf16"
    m.precision = "fp32"
    with pytest.raises(_lib.WdError):
        m.engine()                      # parameters on the CPU: no CPU path in either precision
    cfg = make_config(variant=_lib.VARIANT_PHOSC, in_channels=4, model_channels=320, out_channels=4, num_res_blocks=1,
                      attention_resolutions=(1, 1), channel_mult=(1, 1), num_heads=4, num_head_channels=-1, transformer_depth=1,
                      context_dim=320, vocab_size=53, num_classes=339, max_seq_len=10, latent_hw=(8, 32), add_label_emb=True,
                      phosc_len=769)
    assert (cfg.variant, cfg.n_channel_mult, list(cfg.channel_mult)[:2], cfg.n_attention_resolutions, cfg.phosc_len) == \
        (1, 2, [1, 1], 1, 769)
    assert (cfg.latent_h, cfg.latent_w, cfg.num_classes, cfg.add_label_emb) == (8, 32, 339, 1)
    with pytest.raises(_lib.WdError):
        F32Engine(cfg, (8, 32), "cpu")


def test_attention_maps_variant_state_dict_layout():
    """args.attentionMaps == 1 (unet.py:1336-1364): middle_block1.{0,1}.* instead of middle_block.* -- same keys, order and shapes
    as the reference module built with that flag (spec dumped by oracle/make_golden_attnmaps.py); the engines see the layers
    under the attentionMaps == 0 names, in the same order."""
    import json
    m = UNetModel(args=default_args("cpu", attentionMaps=1), **KW)
    with open(os.path.join(os.path.dirname(__file__), "golden", "state_dict_spec_unet_attnmaps.json")) as f:
        spec = [(k, tuple(sh)) for k, sh in json.load(f)]
    assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == spec
    assert sum(k.startswith("middle_block1.") for k, _ in spec) == 54 and not any(k.startswith("middle_block.") for k, _ in spec)
    assert [k for k, _ in m._engine_state_items()] == [k for k, _ in W.load_spec("unet")]
    assert m.precision == "fp32"      # the maps are the attention probabilities: only the fp32 path materialises them
    m.precision = "bf16"
    x = torch.zeros(1, 4, 8, 32)
    with pytest.raises(NotImplementedError):
        m(x, None, timesteps=torch.tensor([1]), context=torch.zeros(1, 10, dtype=torch.long), y=torch.tensor([0]))


def test_tf32_split_arithmetic_the_tensor_core_fp32_route_relies_on():
    """csrc/f32_gemm_tc.cu: a = hi + lo with hi = tf32_rn(a) (low 13 mantissa bits zero) and lo = a - hi EXACT in fp32; the three-term
    product hi*whi + lo*whi + hi*wlo (each factor seen by the tensor core with at most 10 mantissa bits of lo / wlo kept) is within
    ~2^-20 of a*w, whereas a two-term bf16 split leaves ~2^-16."""
    import numpy as np
    rng = np.random.default_rng(0)
    a = rng.standard_normal(20000).astype(np.float32)
    w = rng.standard_normal(20000).astype(np.float32)

    def rn(x, drop):   # round-to-nearest (ties away, like cvt.rna) to a mantissa with `drop` low bits cleared
        u = x.view(np.uint32).astype(np.uint64)
        u = (u + (1 << (drop - 1))) & ~np.uint64((1 << drop) - 1)
        return u.astype(np.uint32).view(np.float32)

    def split(x, drop):
        hi = rn(x, drop)
        lo = (x - hi).astype(np.float32)
        assert np.array_equal(hi.astype(np.float64) + lo.astype(np.float64), x.astype(np.float64))   # exact
        assert np.all(np.abs(lo) <= np.abs(x) * 2.0 ** -(23 - drop) * 1.0001)
        return hi, rn(lo, drop)     # the hardware keeps only the same mantissa width of lo

    exact = a.astype(np.float64) * w.astype(np.float64)
    errs = {}
    for name, drop in (("tf32", 13), ("bf16", 16)):
        ah, al = split(a, drop)
        wh, wl = split(w, drop)
        approx = (ah.astype(np.float64) * wh + al.astype(np.float64) * wh + ah.astype(np.float64) * wl)
        errs[name] = float(np.max(np.abs(approx - exact) / np.abs(exact)))
    assert errs["tf32"] < 2.0 ** -19 and errs["bf16"] > 2.0 ** -17, errs


def test_phosc_model_refuses_training_mode():
    """in training mode UNetModelPhosc.forward used to run the inference engine and return a tensor without grad_fn
    (a loss.backward() that trains nothing).  It must raise; eval() / no_grad() inference is unaffected (and then fails only
    because this box has no GPU)."""
    from worddiffusion_b200 import _lib
    from worddiffusion_b200.unetPhosc import UNetModelPhosc
    from worddiffusion_b200.unetPhosc2 import UNetModelPhosc as UNetModelPhosc2
    x = torch.zeros(1, 4, 8, 32)
    kw = dict(timesteps=torch.tensor([5]), context=torch.ones(1, 10, dtype=torch.long), y=torch.tensor([1]))
    for cls in (UNetModelPhosc, UNetModelPhosc2):
        m = cls(args=default_args("cpu"), **KW)
        m.train()
        with pytest.raises(NotImplementedError, match="training UNetModelPhosc"):
            m(x, torch.zeros(1, 769), **kw)
        with torch.no_grad(), pytest.raises(_lib.WdError):   # inference call shape is fine; only the device is not
            m(x, torch.zeros(1, 769), **kw)
        m.eval().requires_grad_(False)
        with pytest.raises(_lib.WdError):
            m(x, torch.zeros(1, 769), **kw)


def test_ddim_timesteps_validation():
    d = D.Diffusion(noise_steps=1000, device="cpu")
    assert d.ddim_timesteps(1000) == list(range(999, -1, -1))
    assert d.ddim_timesteps(30)[0] == 957 and len(d.ddim_timesteps(30)) == 30   # "leading" spacing: stride 33
    for bad in (0, -3, 1001):
        with pytest.raises(ValueError):
            d.ddim_timesteps(bad)


def test_parameter_cache_sees_replaced_parameters():
    """The cached parameter walk of `_weights_signature` re-checks object identity on every call."""
    m = UNetModel(args=default_args("cpu"), **KW)
    s1 = m._weights_signature()
    assert m._weights_signature() == s1
    old = m.out[2].bias
    m.out[2].bias = torch.nn.Parameter(old.detach().clone())
    s2 = m._weights_signature()
    assert s2 != s1 and len(s2) == len(s1)
    with torch.no_grad():
        m.out[2].bias.add_(1.0)          # in-place change: version counter
    assert m._weights_signature() != s2
    assert len(s1) == len(list(m.parameters()))


def _empty_shard_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_total = 1                                      # fewer latents than ranks: rank 1 owns nothing
    lo, hi = D.shard_bounds(n_total, world, rank)
    full = torch.arange(n_total * 4 * 8 * 32, dtype=torch.float32).reshape(n_total, 4, 8, 32)
    out = D.all_gather_latents(full[lo:hi].clone(), n_total, world)
    q.put((rank, hi - lo, bool(torch.equal(out, full))))
    dist.destroy_process_group()


def test_all_gather_with_an_empty_shard_world2_gloo():
    """with N < world size the rank without latents must still contribute (a zero-row pad) to the collective."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_empty_shard_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, 1, True), (1, 0, True)]


def test_gradient_bucket_plan():
    """plan_grad_buckets (training.py): contiguous, covering, cut only at stage boundaries, last bucket ends the pass."""
    from worddiffusion_b200.training import plan_grad_buckets
    stages = [0, 0, 1, 2, 2, 3, 5, 5]
    sizes = [64, 64, 192, 64, 64, 256, 64, 64]
    for nb in (1, 2, 3, 4, 8, 50):
        b = plan_grad_buckets(stages, sizes, 7, nb)
        assert 1 <= len(b) <= nb and b[0][1] == 0 and b[-1][2] == sum(sizes) and b[-1][0] == 7
        assert all(x[2] == y[1] for x, y in zip(b, b[1:])) and all(x[0] < y[0] for x, y in zip(b, b[1:]))
        pos = 0
        for st, sz in zip(stages, sizes):  # a slice lies in a bucket whose stage_end is past the slice's final stage
            k = next(i for i, x in enumerate(b) if x[1] <= pos < x[2])
            assert b[k][0] > st and pos + sz <= b[k][2]
            pos += sz
    with pytest.raises(ValueError):
        plan_grad_buckets([1, 0], [64, 64], 2, 2)


def _bucket_worker(rank, world, port, q):
    import torch.distributed as dist
    from worddiffusion_b200.training import plan_grad_buckets
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(100 + rank)
    flat = torch.randn(1024, generator=g)
    whole = flat.clone()
    dist.all_reduce(whole)
    works = [dist.all_reduce(flat[lo:hi], async_op=True) for _, lo, hi in plan_grad_buckets([0, 1, 2, 3], [256] * 4, 4, 3)]
    for w in works:
        w.wait()
    q.put((rank, bool(torch.equal(flat, whole))))
    dist.destroy_process_group()


def test_bucketed_allreduce_equals_whole_allreduce_gloo():
    """world_size 2 on CPU (gloo): all-reducing the buckets of the flat gradient one by one, asynchronously, is the
    all-reduce of the whole buffer."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_bucket_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = dict(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(60)
    assert res == {0: True, 1: True}


def test_subpixel_form_of_upsample_conv():
    """The identity behind GemmArgs::up_phase / f32tc_upconv (csrc/ops.cu upconv_phase_fold_kernel): conv3x3(nearest-2x(x)) ==
    four 2x2 convolutions of x, one per output phase (a, b), whose weights are the sums of the 3x3 taps that read the same
    input pixel (rows a = 0: {0}, {1, 2}; a = 1: {0, 1}, {2}; same for the columns).  fp64, reference ops only."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(0)
    B, C, Co, H, W = 2, 8, 6, 4, 16
    x = torch.randn(B, C, H, W, generator=g, dtype=torch.float64)
    w = torch.randn(Co, C, 3, 3, generator=g, dtype=torch.float64)
    b = torch.randn(Co, generator=g, dtype=torch.float64)
    ref = F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), w, b, padding=1)   # unet.py:497-499
    taps = {0: ([0], [1, 2]), 1: ([0, 1], [2])}
    xp = F.pad(x, (1, 1, 1, 1))
    out = torch.zeros_like(ref)
    for a in (0, 1):
        for bb in (0, 1):
            acc = torch.zeros(B, Co, H, W, dtype=torch.float64)
            for ty in (0, 1):
                for tx in (0, 1):
                    wf = sum(w[:, :, ky, kx] for ky in taps[a][ty] for kx in taps[bb][tx])
                    dy, dx = ty - 1 + a, tx - 1 + bb
                    acc += torch.einsum("oc,bchw->bohw", wf, xp[:, :, 1 + dy:1 + dy + H, 1 + dx:1 + dx + W])
            out[:, :, a::2, bb::2] = acc + b[None, :, None, None]
    assert float((out - ref).abs().max()) < 1e-12
