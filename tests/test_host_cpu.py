"""CPU-side checks: the C-ABI library loads and exports every symbol declared in include/pd_b200.h (no compute
calls), the ctypes struct mirrors the header, host-side schedule code matches the reference's golden scalars,
the checkpoint key grammar matches the reference's, and the product never imports the oracle."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(REPO, "include", "pd_b200.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"#ifdef PD_DEBUG.*?#endif", "", src, flags=re.S)     # PD_DEBUG-only hooks are not in the shipped library
    return sorted(set(re.findall(r"\b(pd_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from prompt_diffusion_b200 import _lib
    syms = _declared_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(_lib.lib, s), f"libpd_b200.so does not export {s}"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature"
    assert _lib.lib.pd_abi_version() == 1
    assert isinstance(_lib.last_error(), str)
    # the shipped library carries no pipeline-sabotaging timing hooks
    assert not hasattr(_lib.lib, "pd_debug_gemm_mode") and not hasattr(_lib.lib, "pd_debug_timeline")


def test_conv_params_struct_matches_header():
    from prompt_diffusion_b200._lib import ConvParams
    src = open(HEADER).read()
    body = src[src.index("typedef struct pd_conv_params {"):src.index("} pd_conv_params;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for line in body.split(";"):
        line = line.strip()
        m = re.match(r"(const\s+)?(void|float|int32_t|int64_t)\s*\*?\s*(.+)$", line.replace("typedef struct pd_conv_params {", "").strip())
        if m:
            names += [n.strip(" *") for n in m.group(3).split(",")]
    assert names == [f[0] for f in ConvParams._fields_]
    # 7 pointers, 19 int32 + float, 2 pointers | round 2: 3 pointers, 4 int64, float + 5 int32 (+4 bytes tail padding)
    assert ctypes.sizeof(ConvParams) == 7 * 8 + 20 * 4 + 2 * 8 + 3 * 8 + 4 * 8 + 6 * 4


def test_schedule_matches_reference_golden(golden):
    from prompt_diffusion_b200.schedule import make_ddim_sampling_parameters, make_ddim_timesteps
    acp = torch.tensor(golden["alphas_cumprod"])
    for S, eta in ((20, 0.0), (50, 0.0), (50, 0.5)):
        ts = make_ddim_timesteps("uniform", S, 1000, verbose=False)
        sig, al, alp = make_ddim_sampling_parameters(acp, ts, eta, verbose=False)
        tag = f"sched_S{S}_eta{eta}"
        assert np.array_equal(ts, golden[tag + "_timesteps"])
        assert np.array_equal(np.asarray(al), golden[tag + "_alphas"])
        assert np.array_equal(np.asarray(alp), golden[tag + "_alphas_prev"])
        assert np.array_equal(np.asarray(sig), golden[tag + "_sigmas"])
        assert np.array_equal(np.asarray(np.sqrt(1. - al)), golden[tag + "_sqrt_one_minus"])


def test_topology_and_key_grammar(cfg):
    from prompt_diffusion_b200.config import CLDMConfig, build_topology
    from prompt_diffusion_b200.synth import param_specs
    specs = param_specs(cfg)
    keys = [s[0] for s in specs]
    assert len(keys) == len(set(keys)) == 1042                      # 686 UNet + 356 ControlNet
    assert sum(k.startswith("model.diffusion_model.") for k in keys) == 686
    n = lambda p: sum(int(np.prod(s[1])) for s in specs if s[0].startswith(p))
    assert n("model.diffusion_model.") == 859_520_964 and n("control_model.") == 362_366_032
    for must in ("control_model.input_hint_block.0.weight", "control_model.input_cond_block.14.weight",
                 "control_model.zero_convs.11.0.weight", "control_model.middle_block_out.0.bias",
                 "model.diffusion_model.input_blocks.4.1.transformer_blocks.0.attn2.to_k.weight",
                 "model.diffusion_model.output_blocks.8.2.conv.weight", "model.diffusion_model.out.2.weight"):
        assert must in keys
    topo = build_topology(cfg)
    assert topo.input_chans == [320, 320, 320, 320, 640, 640, 640, 1280, 1280, 1280, 1280, 1280]
    assert [b[0].cin for b in topo.output_blocks] == [2560, 2560, 2560, 2560, 2560, 1920, 1920, 1280, 960, 960, 640, 640]
    y = CLDMConfig.from_yaml(os.path.join(REPO, "tests", "golden", "cldm_v15_topology.yaml"))
    assert y == cfg


def test_product_does_not_import_oracle():
    pkg = os.path.join(REPO, "prompt-diffusion_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "/root/reference" not in txt, f


def test_model_refuses_cpu():
    from prompt_diffusion_b200 import ControlLDM
    with pytest.raises(RuntimeError):
        ControlLDM(device="cpu")


def test_shard_bounds_cover_batch():
    from prompt_diffusion_b200.parallel import shard_bounds
    for n in (1, 7, 8, 64, 65):
        for world in (1, 2, 4, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["PD_REPO"])
from prompt_diffusion_b200.parallel import sample_sharded
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + os.environ["PD_PORT"],
                        rank=int(os.environ["PD_RANK"]), world_size=int(os.environ["PD_WORLD"]))
B = int(os.environ["PD_BATCH"])
g = torch.Generator().manual_seed(0)
cond = {"c_crossattn": [torch.randn(B, 77, 8, generator=g)], "example_pair": [torch.rand(B, 6, 16, 16, generator=g)],
        "query": [torch.rand(B, 3, 16, 16, generator=g)]}
un = {"c_crossattn": [torch.randn(B, 77, 8, generator=g)], "example_pair": cond["example_pair"], "query": cond["query"]}
x_T = torch.randn(B, 4, 2, 2, generator=g)
def fake_sample(b, c, u, xt):   # per-sample arithmetic only, like the real loop
    assert xt.shape[0] == b == c["query"][0].shape[0] == u["c_crossattn"][0].shape[0]
    return xt * 2 + c["c_crossattn"][0].mean((1, 2))[:, None, None, None] - u["c_crossattn"][0].amax((1, 2))[:, None, None, None] \
        + c["example_pair"][0].sum((1, 2, 3))[:, None, None, None]
full = fake_sample(B, cond, un, x_T)
out = sample_sharded(fake_sample, B, cond, un, x_T, chunk=int(os.environ.get("PD_CHUNK", "0")) or None)
assert out.shape == full.shape and torch.equal(out, full), (out.shape, full.shape)
dist.barrier(); dist.destroy_process_group()
print("ok")
"""


@pytest.mark.parametrize("world,batch,chunk", [(2, 8, 0), (2, 7, 2), (3, 8, 0)])
def test_sharded_sampling_equals_single_process_gloo(world, batch, chunk, tmp_path):
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    procs = []
    for r in range(world):
        env = dict(os.environ, PD_REPO=REPO, PD_PORT=str(port), PD_RANK=str(r), PD_WORLD=str(world),
                   PD_BATCH=str(batch), PD_CHUNK=str(chunk), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, "-c", _WORKER], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT))
    for p in procs:
        out, _ = p.communicate(timeout=180)
        assert p.returncode == 0, out.decode()[-2000:]


# ---- checkpoint ingest (cldm/model.py:8-28, tool_add_control.py:17-48) ---------------------------------------
def test_load_state_dict_reads_ckpt_and_safetensors(tmp_path):
    from prompt_diffusion_b200.cldm.model import get_state_dict, load_state_dict
    sd = {"control_model.zero_convs.0.0.weight": torch.randn(4, 4, 1, 1),
          "model.diffusion_model.out.2.bias": torch.randn(4), "first_stage_model.decoder.conv_in.weight": torch.randn(2, 2)}
    assert get_state_dict({"state_dict": sd}) is sd and get_state_dict(sd) is sd
    p1 = str(tmp_path / "a.ckpt")
    torch.save({"state_dict": sd, "global_step": 7}, p1)            # Lightning envelope
    p2 = str(tmp_path / "b.pth")
    torch.save(sd, p2)                                              # bare dict
    import safetensors.torch
    p3 = str(tmp_path / "c.safetensors")
    safetensors.torch.save_file(sd, p3)
    for p in (p1, p2, p3):
        got = load_state_dict(p, location="cpu")
        assert set(got) == set(sd)
        for k in sd:
            assert torch.equal(got[k], sd[k])


def test_add_control_key_scheme():
    """control_<x> copies model.diffusion_<x>; other keys copy themselves; missing ones keep the scratch init."""
    from prompt_diffusion_b200.cldm.model import add_control, get_node_name, path_keys
    assert get_node_name("control_model.a", "control_") == (True, "model.a")
    assert get_node_name("control_", "control_") == (False, "")
    assert get_node_name("model.diffusion_model.a", "control_") == (False, "")
    pre = {"state_dict": {"model.diffusion_model.input_blocks.1.0.in_layers.2.weight": torch.full((2, 2), 3.0),
                          "model.diffusion_model.out.2.weight": torch.full((2,), 5.0),
                          "first_stage_model.x": torch.ones(1)}}
    scratch = {"control_model.input_blocks.1.0.in_layers.2.weight": torch.zeros(2, 2),
               "control_model.zero_convs.0.0.weight": torch.full((1,), 9.0),
               "model.diffusion_model.input_blocks.1.0.in_layers.2.weight": torch.zeros(2, 2),
               "model.diffusion_model.out.2.weight": torch.zeros(2),
               "first_stage_model.x": torch.zeros(1)}
    tgt, added = add_control(pre, scratch, verbose=False)
    assert set(tgt) == set(scratch) and added == ["control_model.zero_convs.0.0.weight"]
    assert float(tgt["control_model.input_blocks.1.0.in_layers.2.weight"][0, 0]) == 3.0      # copied from the UNet
    assert float(tgt["model.diffusion_model.out.2.weight"][0]) == 5.0
    assert float(tgt["control_model.zero_convs.0.0.weight"][0]) == 9.0                       # scratch value kept
    assert tgt["first_stage_model.x"].data_ptr() != pre["state_dict"]["first_stage_model.x"].data_ptr()   # clones
    assert sorted(path_keys(tgt)) == sorted(k for k in scratch if not k.startswith("first_stage"))


def test_create_model_refuses_cpu():
    from prompt_diffusion_b200.cldm.model import create_model
    with pytest.raises(RuntimeError):
        create_model(os.path.join(REPO, "tests", "golden", "cldm_v15_topology.yaml"), device="cpu")


def test_ddim_scheduler_restatement_matches_reference_schedule(golden):
    """The restated diffusers DDIMScheduler (SD1.5 config) must walk the reference's own DDIM schedule
    (cldm/ddim_hacked.py make_schedule, golden sched_S*_eta*): timesteps 981 ... 1, alpha_t, alpha_prev (alpha_cumprod[0]
    after the last step), sigma — and its step() must equal ddim_hacked.py:218-233 on random tensors."""
    from prompt_diffusion_b200.pipeline_prompt_diffusion import DDIMScheduler
    for S, eta in ((20, 0.0), (50, 0.0), (50, 0.5)):
        tag = f"sched_S{S}_eta{eta}"
        sch = DDIMScheduler()
        sch.set_timesteps(S)
        ts = sch.timesteps.numpy()
        assert np.array_equal(ts[::-1], golden[tag + "_timesteps"])
        g = torch.Generator().manual_seed(S)
        x, e = torch.randn(2, 4, 8, 8, generator=g), torch.randn(2, 4, 8, 8, generator=g)
        for i, t in enumerate(ts):
            index = S - 1 - i
            a_t, a_prev = float(golden[tag + "_alphas"][index]), float(golden[tag + "_alphas_prev"][index])
            sigma, s1m = float(golden[tag + "_sigmas"][index]), float(golden[tag + "_sqrt_one_minus"][index])
            assert abs(float(sch.alphas_cumprod[t]) - a_t) <= 2e-6 * a_t      # diffusers builds the betas in fp32, ldm in fp64
            pred_x0 = (x - s1m * e) / a_t ** 0.5
            ref = a_prev ** 0.5 * pred_x0 + (1.0 - a_prev - sigma ** 2) ** 0.5 * e          # noise term checked via sigma below
            got = sch.step(e, int(t), x, eta=0.0)[0] if eta == 0.0 else None
            if got is not None:
                assert float((got - ref).abs().max()) <= 5e-5 * float(ref.abs().max())
            else:
                prev_t = int(t) - 1000 // S
                ap = float(sch.alphas_cumprod[prev_t]) if prev_t >= 0 else float(sch.final_alpha_cumprod)
                var = (1 - ap) / (1 - a_t) * (1 - a_t / ap)
                assert abs(eta * var ** 0.5 - sigma) <= 1e-5 * max(sigma, 1e-6) + 1e-9 and abs(ap - a_prev) <= 2e-6 * a_prev


def test_ddim_scheduler_step_coefficients_feed_the_fused_step(golden):
    """``DDIMScheduler.step_coefficients`` — the rows the pipeline's fused route hands to ``pd_cfg_ddim_step`` — are the
    reference schedule's own scalars (cldm/ddim_hacked.py:206-214, golden sched_S*_eta*) and reproduce ``step()``."""
    from prompt_diffusion_b200.pipeline_prompt_diffusion import DDIMScheduler
    for S, eta in ((20, 0.0), (50, 0.0), (50, 0.5)):
        tag = f"sched_S{S}_eta{eta}"
        sch = DDIMScheduler()
        sch.set_timesteps(S)
        rows = sch.step_coefficients(eta)
        assert len(rows) == S
        g = torch.Generator().manual_seed(100 + S)
        x, e = torch.randn(2, 4, 8, 8, generator=g), torch.randn(2, 4, 8, 8, generator=g)
        for i, (t, (a_t, a_prev, sigma, s1m)) in enumerate(zip(sch.timesteps.tolist(), rows)):
            index = S - 1 - i
            assert abs(a_t - float(golden[tag + "_alphas"][index])) <= 2e-6 * a_t
            assert abs(a_prev - float(golden[tag + "_alphas_prev"][index])) <= 2e-6 * a_prev
            assert abs(sigma - float(golden[tag + "_sigmas"][index])) <= 1e-5 * max(sigma, 1e-3)
            assert abs(s1m - float(golden[tag + "_sqrt_one_minus"][index])) <= 2e-6
            if eta == 0.0:
                fused = a_prev ** 0.5 * ((x - s1m * e) / a_t ** 0.5) + (1.0 - a_prev - sigma ** 2) ** 0.5 * e
                got = sch.step(e, int(t), x, eta=0.0)[0]
                assert float((got - fused).abs().max()) <= 1e-5 * float(fused.abs().max())


def test_diffusers_controlnet_refuses_configs_outside_the_path():
    """Class / additional embeddings (promptdiffusioncontrolnet.py:288-320) are not part of the SD1.5 prompt-diffusion
    configuration: a config that asks for them is refused at construction, before any device is touched."""
    from prompt_diffusion_b200.promptdiffusioncontrolnet import PromptDiffusionControlNetModel
    for kw in ({"class_embed_type": "timestep"}, {"num_class_embeds": 10}, {"addition_embed_type": "text_time"}):
        with pytest.raises(NotImplementedError):
            PromptDiffusionControlNetModel(device="cpu", **kw)
