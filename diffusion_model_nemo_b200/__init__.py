"""B200-native reverse-diffusion sampling path: drop-in for `diffusion_model_nemo`'s sampler / U-Net API.

    from diffusion_model_nemo_b200.modules import Unet, GaussianDiffusion
"""
__version__ = "0.1.0"
