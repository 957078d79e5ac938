"""B200-native MultiDimStacker sliding-window forward path (drop-in for lRomul/ball-action-spotting's
``src/models/multidim_stacker.py`` and ``src/predictors.py``).  See DESIGN.md / INTEGRATION.md."""
from .indexes import StackIndexesGenerator  # noqa: F401
from .frames import get_frames_processor, PadNormalizeFramesProcessor  # noqa: F401
from .model import MultiDimStacker  # noqa: F401
from .predictor import MultiDimStackerPredictor, load_model  # noqa: F401
from .train import FrozenEncoderTrainer  # noqa: F401

nn_module_registry = {"multidim_stacker": MultiDimStacker}   # plug point of BallActionModel.nn_module (argus_models.py:18-21)
