"""Same export surface as the reference's `diffusion_model_nemo.modules` for the sampling hot path, so a Hydra
`_target_: diffusion_model_nemo.modules.X` becomes `_target_: diffusion_model_nemo_b200.modules.X`."""
from .unet import Unet, WaveGradUNet
from .diffusion_process import (
    linear_beta_schedule,
    quadratic_beta_schedule,
    cosine_beta_schedule,
    sigmoid_beta_schedule,
    AbstractDiffusionProcess,
    CosineSchedule,
    LinearSchedule,
    QuadraticSchedule,
    SigmoidSchedule,
)
from .gaussian_diffusion import GaussianDiffusion
from .learned_gaussian_diffusion import LearnedGaussianDiffusion
from .generalized_gaussian_diffusion import GeneralizedGaussianDiffusion
from .wavegrad_diffusion import WaveGradDiffusion
from .sde import (
    SDE,
    VPSDE,
    VESDE,
    Predictor,
    NonePredictor,
    EulerMaruyamaPredictor,
    AncestralSamplingPredictor,
    ReverseDiffusionPredictor,
    register_predictor,
    get_predictor,
    Corrector,
    NoneCorrector,
    LangevinCorrector,
    AnnealedLangevinDynamics,
    get_corrector,
    register_corrector,
    PredictorCorrectorSampler,
    resolve_score_function,
)
