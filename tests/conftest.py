import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with `pytest -m gpu`")
    # The tests build dozens of U-Nets: they run the STATIC "mixed" profile (unet_engine.MIXED_PROFILES) instead of paying the load-time
    # precision calibration (upgpt_b200/precision.py, ~3 s per model) each time; tests/test_gpu_hotpath.py::test_precision_calibration*
    # and the benchmarked-configuration parity test switch it back on.
    os.environ.setdefault("UPGPT_CALIBRATE", "0")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "hotpath_golden.npz"))


def tiny_ldm_config():
    """A small LatentDiffusion config with the same structure as configs/deepfashion/bbox.yaml (hybrid conditioning,
    text + style + SMPL context, KL VAE), sized for fast tests."""
    from oracle.make_golden import TINY_UNET_KW, TINY_VAE_KW
    return {
        "target": "ldm.models.diffusion.ddpm.LatentDiffusion",
        "params": {
            "linear_start": 0.00085, "linear_end": 0.012, "num_timesteps_cond": 1, "log_every_t": 1000, "timesteps": 1000,
            "first_stage_key": "image", "cond_stage_key": "txt", "concat_key": "person_mask", "image_size": [16, 16],
            "channels": 4, "cond_stage_trainable": False, "conditioning_key": "hybrid", "scale_factor": 0.18215, "use_ema": False,
            "unet_config": {"target": "ldm.modules.diffusionmodules.openaimodel.UNetModel", "params": dict(TINY_UNET_KW)},
            "first_stage_config": {"target": "ldm.models.autoencoder.AutoencoderKL",
                                   "params": {"embed_dim": 4, "ddconfig": dict(TINY_VAE_KW), "lossconfig": {"target": "torch.nn.Identity"}}},
            "cond_stage_config": {"target": "ldm.modules.encoders.modules.FrozenCLIPEmbedder"},
            "extra_cond_stages": {
                "style_cond": {"target": "ldm.modules.poses.poses.DummyModel", "cond_stage_key": "styles"},
                "pose_cond": {"target": "ldm.modules.poses.poses.LinearProject", "cond_stage_key": "smpl",
                              "params": {"input_dim": 85, "output_dim": 128}}},
        },
    }
