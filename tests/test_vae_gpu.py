"""First-stage decode on the CUDA path against the reference's own Decoder output
(tests/golden/vae_decoder_golden.npz, made by tests/golden/make_golden_vae.py) and against the oracle at the
BASELINE latent size.  Gates: image rel-L2 <= 1e-4 in fp32 mode, <= 2e-2 in bf16 mode (measured 1.5e-2 at every size:
the decoder is ~30 3x3 convolutions in series on bf16 operands with bf16 block outputs; its image is consumed at
8 bits, i.e. at 4e-3 of full range)."""
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = {"fp32": 1e-4, "bf16": 2e-2}


@pytest.fixture(scope="module", autouse=True)
def _no_tf32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.set_grad_enabled(False)
    yield
    torch.set_grad_enabled(True)


@pytest.fixture(scope="module")
def decoders(vae_state_dict_cpu, golden_vae):
    from prompt_diffusion_b200 import AutoencoderKLDecoder
    sf = float(golden_vae["scale_factor"])
    return {m: AutoencoderKLDecoder(m, DEV).load_state_dict(vae_state_dict_cpu, scale_factor=sf) for m in ("fp32", "bf16")}


@pytest.mark.parametrize("rows,cols", [(64, 4096), (37, 256), (5, 16384), (3, 200)])
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_softmax_rows(rows, cols, dt):
    """pd_softmax_rows vs torch (fp32 statistics) on the operand's own dtype, in place and pitched."""
    from prompt_diffusion_b200 import ops
    if dt == torch.float32 and cols > 8192:
        pytest.skip("fp32 rows are limited to 8192 columns")
    g = torch.Generator(device=DEV).manual_seed(5)
    full = (torch.randn(rows, cols + 16, device=DEV, generator=g) * 3).to(dt)
    x = full[:, :cols]
    ref = torch.softmax(x.float() * 0.37, dim=-1)
    out = torch.empty(rows, cols, device=DEV, dtype=dt)
    ops.softmax_rows(x, out, 0.37)
    assert rel_l2(out.float(), ref) < (2e-6 if dt == torch.float32 else 4e-3)
    ops.softmax_rows(x, x, 0.37)                                   # in place, pitched
    assert torch.equal(full[:, :cols], out)
    assert bool((full[:, cols:] != 0).any())                       # the pad columns were not touched (still random)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("name", ["z16", "z8x24"])
def test_decode_first_stage_vs_reference_golden(decoders, golden_vae, mode, name):
    z = torch.tensor(golden_vae[name + "_z"], device=DEV)
    img = decoders[mode].decode(z, scaled=True)
    ref = torch.tensor(golden_vae[name + "_img"])
    assert img.shape == ref.shape and img.dtype == torch.float32
    err = rel_l2(img.cpu(), ref)
    print(f"[parity] vae decode {name} {mode}: rel-L2 = {err:.3e}")
    assert err <= TOL[mode], (name, mode, err)


def test_decode_unscaled_matches_autoencoder_decode(decoders, golden_vae, vae_state_dict_cpu):
    """AutoencoderKL.decode (no latent scaling) vs the oracle."""
    from oracle import vae_oracle as V
    z = torch.tensor(golden_vae["z16_z"], device=DEV)
    ref = V.decode({k: v.to(DEV) for k, v in vae_state_dict_cpu.items()}, z)
    assert rel_l2(decoders["fp32"].decode(z).cpu(), ref.cpu()) <= TOL["fp32"]


def test_decode_config2_latent_vs_oracle_gpu(decoders, vae_state_dict_cpu, golden_vae):
    """64x64 latent (512^2 image, 4096-token middle attention), batch 2: CUDA path vs the oracle run on the GPU in fp32."""
    from oracle import vae_oracle as V
    g = torch.Generator(device=DEV).manual_seed(3)
    z = torch.randn(2, 4, 64, 64, device=DEV, generator=g)
    sd = {k: v.to(DEV) for k, v in vae_state_dict_cpu.items()}
    ref = V.decode_first_stage(sd, z, float(golden_vae["scale_factor"]))
    for mode in ("fp32", "bf16"):
        img = decoders[mode].decode(z, scaled=True)
        err = rel_l2(img.cpu(), ref.cpu())
        print(f"[parity] vae decode 512^2 {mode} vs oracle(gpu fp32): rel-L2 = {err:.3e}")
        assert img.shape == (2, 3, 512, 512) and err <= TOL[mode], (mode, err)
        decoders[mode].release_buffers()


def test_controlldm_decode_first_stage_hook(vae_state_dict_cpu, golden_vae, cfg):
    """ControlLDM.decode_first_stage exists, refuses without weights, and AutoencoderKLDecoder refuses the CPU."""
    from prompt_diffusion_b200 import AutoencoderKLDecoder, ControlLDM
    m = ControlLDM(cfg, mode="bf16", device=DEV)
    with pytest.raises(RuntimeError):
        m.decode_first_stage(torch.zeros(1, 4, 8, 8, device=DEV))
    with pytest.raises(RuntimeError):
        AutoencoderKLDecoder("bf16", "cpu")
    assert abs(m.scale_factor - float(golden_vae["scale_factor"])) < 1e-6
