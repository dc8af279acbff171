"""CPU: host logic, C-ABI surface, fail-loudly behaviour, shard partition and 2-rank gloo gather."""
import ctypes
import os
import re
import socket
import sys

import numpy as np
import pytest
import torch

import __graft_entry__ as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def built():
    G.build()


def test_header_symbols_exported():
    import super_diff_disease_b200 as S
    hdr = open(os.path.join(ROOT, "include", "sdd_b200.h")).read()
    declared = set(re.findall(r"\b(sdd_[a-z0-9_]+)\s*\(", hdr))
    L = ctypes.CDLL(S.lib_path())
    for name in declared:
        assert hasattr(L, name), f"{name} declared in sdd_b200.h but not exported"
    assert declared == set(S._lib.SYMBOLS), (declared ^ set(S._lib.SYMBOLS))
    assert S.lib().sdd_abi_version() == 2


def test_no_cpu_fallback():
    import super_diff_disease_b200 as S
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert S.lib().sdd_device_check() != 0
    m = S.UNet()
    with pytest.raises(S.SddError):
        m(torch.zeros(1, 1, 16, 16), torch.zeros(1, dtype=torch.long))
    with pytest.raises(S.SddError):
        S.DDPM(4).sample(m, (1, 1, 16, 16), "cpu")
    with pytest.raises(S.SddError):
        S.superposed_sample([m, m], S.DDPM(4), (1, 1, 16, 16), "cpu", seed=0)
    with pytest.raises(S.SddError):
        S.superposed_sample([m, m], S.DDPM(4), (1, 1, 16, 16), "cpu", seed=0, mode="and")
    with pytest.raises(S.SddError):
        S.superposed_sample([m, m], S.DDPM(4), (1, 1, 16, 16), "cuda", seed=0, mode="xor")  # rejected before any GPU use
    # the "next" rows have no CPU path either: N4 forward pieces, attention core / block
    x0 = torch.zeros(2, 1, 16, 16)
    with pytest.raises(S.SddError):
        S.DDPM(4).q_sample(x0, torch.zeros(2, dtype=torch.long), torch.zeros_like(x0))
    with pytest.raises(S.SddError):
        S.DDPM(4).p_losses(m, x0, torch.zeros(2, dtype=torch.long))
    q = torch.zeros(1, 1, 128, 64, dtype=torch.float16)
    with pytest.raises(S.SddError):
        S.attention_core(q, q, q)
    with pytest.raises(S.SddError):
        S.attention_block(torch.zeros(1, 16, 16, 128, dtype=torch.float16), torch.ones(128), torch.zeros(128),
                          torch.zeros(384, 128), torch.zeros(384), torch.zeros(128, 128), torch.zeros(128))
    # the C ABI itself refuses without an sm_100 device (every compute entry point starts with the device check)
    L = S.lib()
    assert L.sdd_attention_fwd(None, None, None, None, 1, 128, 64, 0.125, None) != 0
    assert L.sdd_q_sample(None, None, None, None, None, 1, 4, None) != 0


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "super-diff-disease_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_state_dict_contract():
    """Same 54 keys / shapes as the reference UNet (enumerated in SURVEY section 5)."""
    import super_diff_disease_b200 as S
    from oracle.superdiff_oracle import init_unet_params
    p = init_unet_params(0)
    m = S.UNet()
    sd = m.state_dict()
    assert list(sd.keys()) == list(p.keys())
    for k in sd:
        assert sd[k].shape == p[k].shape, k
    m.load_state_dict(p, strict=True)
    with pytest.raises(S.SddError):
        S.UNet(base_channels=32)


def test_unet_copies_never_share_or_pickle_the_c_handle(tmp_path):
    """ema_pytorch deep-copies the model and torch.save(model) pickles it (training_logic.py:16,47-48,55).  The C handle
    is a ctypes pointer (unpicklable; a cloned one would be freed twice): copies drop it and rebuild lazily."""
    import copy
    import pickle
    import super_diff_disease_b200 as S
    m = S.UNet()
    m._handle, m._handle_key = ctypes.c_void_p(0xDEAD0), ("fake",)  # as if a forward had run
    c = copy.deepcopy(m)
    assert c._handle is None and c._handle_key is None and m._handle.value == 0xDEAD0
    assert all(torch.equal(a, b) and a.data_ptr() != b.data_ptr() for a, b in zip(m.state_dict().values(), c.state_dict().values()))
    r = pickle.loads(pickle.dumps(m))
    assert r._handle is None and list(r.state_dict()) == list(m.state_dict())
    torch.save(m, tmp_path / "m.pt")
    assert torch.load(tmp_path / "m.pt", weights_only=False)._handle is None
    m._handle = m._handle_key = None  # nothing real to free


def test_training_calls_are_refused_by_name():
    """the swap applies at sampling / evaluation sites only.  In train() mode with gradients
    enabled (training_logic.py:28-36) forward / training_step raise a named error rather than return a loss without a
    grad_fn; the check runs before any device work, so it is testable without a GPU."""
    import super_diff_disease_b200 as S
    m = S.UNet().train()
    x = torch.zeros(2, 1, 16, 16)
    with pytest.raises(S.SddError, match="forward-only"):
        m(x, torch.zeros(2, dtype=torch.long))
    with pytest.raises(S.SddError, match="forward-only"):
        S.DDPM(4).training_step(m, x)
    with pytest.raises(S.SddError, match="forward-only"):
        S.DDPM(4).p_losses(m, x, torch.zeros(2, dtype=torch.long))


def test_reference_modules_installed_and_oracle_matches_them():
    """baseline/_ref (git-ignored copy of the reference's unet.py / ddpm.py, baseline/install_ref.py) loads, and the
    oracle reproduces the reference module's forward bit for bit on CPU; skipped where neither /root/reference nor an
    installed baseline/_ref exists."""
    from baseline import install_ref, ref_loader
    from oracle import superdiff_oracle as O
    install_ref.install(verbose=False)
    ref = ref_loader.load()
    if ref is None:
        pytest.skip("no /root/reference and no baseline/_ref here")
    RefUNet, RefDDPM = ref
    p = O.init_unet_params(1)
    net = RefUNet()
    net.load_state_dict(p, strict=True)
    x = torch.randn(2, 1, 32, 32, generator=torch.Generator().manual_seed(0))
    t = torch.tensor([3, 700])
    with torch.no_grad():
        assert torch.equal(net.eval()(x, t), O.unet_forward(p, x, t))
    assert torch.equal(RefDDPM(num_timesteps=100).alpha_bars, O.Schedule(100).alpha_bars)


def test_ddpm_schedule_matches_golden(golden_dir):
    import super_diff_disease_b200 as S
    g = np.load(os.path.join(golden_dir, "ddpm_sample.npz"))
    d = S.DDPM(1000)
    assert d.T == 1000
    assert np.array_equal(d.betas.numpy(), g["sched1000_betas"])
    assert np.array_equal(d.alpha_bars.numpy(), g["sched1000_alpha_bars"])
    assert torch.equal(d.alphas, 1.0 - d.betas)


def test_shard_range():
    from super_diff_disease_b200 import shard_range
    for gb in (1, 7, 8, 64, 33):
        for ws in (1, 2, 4, 8):
            spans = [shard_range(gb, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == gb
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, gb, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from oracle.superdiff_oracle import philox_normal
    from super_diff_disease_b200 import sharded_sample
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)

    def local_fn(lo, hi):  # stand-in sampler with the product's global-sample-id noise keying
        return torch.from_numpy(philox_normal(99, np.arange(lo, hi), 0, 64)).reshape(hi - lo, 1, 8, 8)

    full = sharded_sample(local_fn, gb, (1, 8, 8), "cpu")

    def local_traj(lo, hi):  # x plus kappa [T,b,M] / log q [T+1,b,M] trajectories keyed by global sample id
        ids = torch.arange(lo, hi, dtype=torch.float32)
        kap = ids[None, :, None] + torch.arange(3.0)[:, None, None] * 100 + torch.arange(2.0)[None, None, :] * 0.5
        lq = ids[None, :, None] - torch.arange(4.0)[:, None, None] * 10 + torch.arange(2.0)[None, None, :] * 0.25
        return local_fn(lo, hi), kap, lq

    x2, kap, lq = sharded_sample(local_traj, gb, (1, 8, 8), "cpu", trajectories=True)
    assert torch.equal(x2, full)
    q.put((rank, full.numpy(), kap.numpy(), lq.numpy()))
    dist.destroy_process_group()


@pytest.mark.parametrize("gb", [4, 5])
def test_two_rank_gloo_gather_is_shard_invariant(gb):
    import torch.multiprocessing as mp
    from oracle.superdiff_oracle import philox_normal
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, gb, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(60)
    res = {r: x for r, x, _, _ in got}
    single = philox_normal(99, np.arange(gb), 0, 64).reshape(gb, 1, 8, 8)
    assert np.array_equal(res[0], single) and np.array_equal(res[1], single)
    # kappa [T, B, M] / log q [T+1, B, M] come back whole, in global sample order, on every rank
    ids = np.arange(gb, dtype=np.float32)
    kap_want = ids[None, :, None] + np.arange(3, dtype=np.float32)[:, None, None] * 100 + np.arange(2, dtype=np.float32)[None, None, :] * 0.5
    lq_want = ids[None, :, None] - np.arange(4, dtype=np.float32)[:, None, None] * 10 + np.arange(2, dtype=np.float32)[None, None, :] * 0.25
    for _, _, kap, lq in got:
        assert np.array_equal(kap, kap_want) and np.array_equal(lq, lq_want)


def test_cli_checkpoint_layout_and_grid(tmp_path):
    """N1: the CLI resolves checkpoints with the reference's layout (env.py:26, training_logic.py:47-48), accepts plain
    and ema_pytorch-prefixed state_dicts, and fails like the reference (FileNotFoundError) on a missing file."""
    import numpy as np
    import torch
    from super_diff_disease_b200 import cli
    assert cli.checkpoint_path("/ck", "exp1", "runA", "TB", 7) == "/ck/exp1/runA/TB/ema_epoch7.pt"
    assert cli.checkpoint_path("/ck", "exp1", "runA", "PNEUMONIA", 7, ema=False) == "/ck/exp1/runA/PNEUMONIA/ddpm_epoch7.pt"
    with pytest.raises(FileNotFoundError):
        cli.main(["--tb", str(tmp_path / "a.pt"), "--pneumonia", str(tmp_path / "b.pt")])
    x = torch.arange(2 * 16 * 8, dtype=torch.float32).reshape(2, 1, 16, 8)
    cli.save_grid_pgm(x, str(tmp_path / "g.pgm"), cols=2)
    raw = open(tmp_path / "g.pgm", "rb").read()
    assert raw.startswith(b"P5\n16 16\n255\n") and len(raw) == len(b"P5\n16 16\n255\n") + 256
    img = np.frombuffer(raw[-256:], dtype=np.uint8).reshape(16, 16)
    assert img[0, 0] == 0 and img[15, 7] == 255 and img[0, 8] == 0 and img[15, 15] == 255


def test_hot_kernels_are_tcgen05_tma_sass():
    """The built library's hot kernels really are Blackwell tensor-core / TMA code: the SASS of the product conv and the
    attention kernel contains UTCHMMA (tcgen05.mma), UTMALDG (TMA tensor loads), LDTM (tcgen05.ld from TMEM) and UTCBAR
    (tcgen05.commit); checked with cuobjdump, no GPU needed."""
    import shutil
    import subprocess
    import collections
    import re
    import __graft_entry__ as G
    lib = G.build()
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", lib], capture_output=True, text=True, check=True).stdout
    cur, cnt = None, collections.defaultdict(collections.Counter)
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            for key in ("UTCHMMA", "UTMALDG", "LDTM", "UTCBAR"):
                if m.group(1).startswith(key):
                    cnt[cur][key] += 1
    conv = [f for f in cnt if "conv3x3_tc4_kernel" in f]
    attn = [f for f in cnt if "attention_fwd_kernel" in f]
    assert len(conv) >= 4 and len(attn) == 1
    for f in conv + attn:
        for key in ("UTCHMMA", "UTMALDG", "LDTM", "UTCBAR"):
            assert cnt[f][key] > 0, (f, key)


def test_bench_config_helpers():
    """bench.py host logic: the metric names the configuration actually run (it was hard-coded), strong
    scaling splits BASELINE configs[2]'s global batch 64 into 64/32/16/8 per GPU, weak scaling keeps it per GPU."""
    import argparse
    import bench
    a = argparse.Namespace(res=512, diffusion_steps=1000, batch=32, scaling="strong", arch="ref")
    assert bench.metric_name(a) == "superposed samples/sec at 512^2 (2 UNets, 1000 steps)"
    assert [bench.per_gpu_batch(a, w) for w in (1, 2, 4, 8)] == [32, 16, 8, 4]
    cfg = bench.workload_config(a, 8)
    assert cfg["global_batch"] == 32 and cfg["per_gpu_batch"] == 4 and "BASELINE configs[3]" in cfg["workload"]
    c3 = argparse.Namespace(res=256, diffusion_steps=250, batch=64, scaling="strong", arch="ref")
    assert [bench.per_gpu_batch(c3, w) for w in (1, 2, 4, 8)] == [64, 32, 16, 8]
    assert "BASELINE configs[2]" in bench.workload_config(c3, 4)["workload"]
    c3.scaling = "weak"
    assert bench.per_gpu_batch(c3, 8) == 64 and bench.workload_config(c3, 8)["global_batch"] == 512
    with pytest.raises(SystemExit):
        bench.per_gpu_batch(argparse.Namespace(batch=10, scaling="strong"), 4)
    ext = argparse.Namespace(res=256, diffusion_steps=250, batch=64, scaling="strong", arch="attn")
    assert "EXTENSION" in bench.metric_name(ext) and "oracle/unet_attn_oracle.py" in bench.workload_config(ext, 1)["workload"]
