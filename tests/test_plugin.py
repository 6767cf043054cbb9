"""AllenAct plugin surface (embclip_b200/plugin.py): the contract of clip_preprocessors.ClipResNetPreprocessor
 -- constructor arguments, attributes, observation space, lazy build, no CPU fallback --
and, on the GPU, that process() equals the oracle on the same frames."""
import pytest
import torch

from conftest import synthetic_frames
from embclip_b200.plugin import ClipResNetPreprocessor, Preprocessor


def test_surface_matches_allenact_contract():
    p = ClipResNetPreprocessor(rgb_input_uuid="rgb_lowres", clip_model_type="RN50", pool=False,
                               output_uuid="rgb_clip_resnet")
    assert isinstance(p, Preprocessor) or type(p).__mro__[1].__name__ == "Preprocessor"
    assert p.input_uuids == ["rgb_lowres"] and p.uuid == "rgb_clip_resnet"
    assert tuple(p.observation_space.shape) == (2048, 7, 7)
    assert p.CLIP_RGB_MEANS == (0.48145466, 0.4578275, 0.40821073)
    assert p.CLIP_RGB_STDS == (0.26862954, 0.26130258, 0.27577711)
    assert p.device == torch.device("cpu") and p._resnet is None          # lazy, CPU until .to()
    pooled = ClipResNetPreprocessor("rgb", "RN50", pool=True)
    assert tuple(pooled.observation_space.shape) == (2048,)
    with pytest.raises(AssertionError):
        ClipResNetPreprocessor("rgb", "RN50x64", pool=False)


def test_no_cpu_fallback(built_lib):
    p = ClipResNetPreprocessor("rgb", "RN50", pool=False)
    with pytest.raises(RuntimeError, match="no CPU path"):
        p.process({"rgb": torch.zeros(1, 224, 224, 3)})


def test_missing_weights_is_an_error(built_lib, monkeypatch):
    monkeypatch.delenv("EMBCLIP_CLIP_WEIGHTS", raising=False)
    monkeypatch.delenv("EMBCLIP_SYNTHETIC_WEIGHTS", raising=False)
    from embclip_b200.plugin import load_clip_visual_state_dict
    with pytest.raises(RuntimeError, match="no weights"):
        load_clip_visual_state_dict("RN50")


@pytest.mark.gpu
@pytest.mark.parametrize("pool", [False, True])
def test_process_matches_oracle(built_lib, rn50_visual, pool):
    assert torch.cuda.is_available()
    sd = {"visual." + k: v for k, v in rn50_visual.state_dict().items()}   # official checkpoint key style
    p = ClipResNetPreprocessor("rgb", "RN50", pool=pool, clip_state_dict=sd).to(torch.device("cuda:0"))
    frames = synthetic_frames(3, seed=9)
    out = p.process({"rgb": frames})                                        # host tensor in, like batch_observations
    assert out.dtype == torch.float32 and out.device.type == "cuda"
    assert tuple(out.shape[1:]) == tuple(p.observation_space.shape)
    with torch.no_grad():
        ref = rn50_visual.trunk(frames.permute(0, 3, 1, 2).contiguous())
        if pool:
            ref = torch.flatten(torch.nn.functional.adaptive_avg_pool2d(ref, (1, 1)), 1)
    err = ((out.cpu() - ref).flatten(1).norm(dim=1) / ref.flatten(1).norm(dim=1)).max().item()
    assert err <= 1e-3, err


@pytest.mark.gpu
def test_depth_input_is_repeated(built_lib, rn50_visual):
    p = ClipResNetPreprocessor("depth", "RN50", pool=True, clip_state_dict=rn50_visual.state_dict()).to("cuda:0")
    d = synthetic_frames(2, seed=2)[..., :1].contiguous()
    a = p.process({"depth": d})
    b = p.process({"depth": d.repeat(1, 1, 1, 3)})
    assert torch.equal(a, b)


def test_vit_surface():
    from embclip_b200.plugin import ClipViTPreprocessor
    p = ClipViTPreprocessor("rgb_lowres", "ViT-B/32")
    assert p.input_uuids == ["rgb_lowres"] and p.uuid == "rgb_clip_vit" and tuple(p.observation_space.shape) == (512,)
    with pytest.raises(AssertionError):
        ClipViTPreprocessor("rgb", "ViT-L/14")
    with pytest.raises(RuntimeError, match="no CPU path"):
        p.process({"rgb_lowres": torch.zeros(1, 224, 224, 3)})


@pytest.mark.gpu
def test_vit_process_matches_oracle(built_lib):
    from embclip_b200.plugin import ClipViTPreprocessor
    from oracle.clip_model import build_vit_b32, freeze_model, init_synthetic_transformer
    torch.manual_seed(0)
    m = freeze_model(init_synthetic_transformer(build_vit_b32(), seed=1234))
    p = ClipViTPreprocessor("rgb", "ViT-B/32", clip_state_dict=m.state_dict()).to(torch.device("cuda:0"))
    frames = synthetic_frames(2, seed=4)
    out = p.process({"rgb": frames})
    with torch.no_grad():
        ref = m.encode_image(frames.permute(0, 3, 1, 2).contiguous())
    err = ((out.cpu() - ref).norm(dim=1) / ref.norm(dim=1)).max().item()
    assert out.shape == (2, 512) and err <= 1e-3, err
