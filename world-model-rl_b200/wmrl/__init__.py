"""wmrl - B200-native RSSM latent-imagination path behind the module API of
mauicv/world-model-rl (package ``reflect``).  See DESIGN.md / INTEGRATION.md."""
from .actor import Actor
from .memory_actor import WorldModelActor
from .models import DenseModel, ValueCritic, set_precision
from .rssm import ContinuousRSSM, DiscreteRSSM, RSSMBase
from .state import (InternalStateContinuous, InternalStateContinuousSequence, InternalStateDiscrete,
                    InternalStateDiscreteSequence)
from .utils import AdamOptim, FreezeParameters
from .value_trainer import ValueGradTrainer, ValueGradTrainerLosses
from .world_model import ImaginedRollouts, WorldModel, WorldModelTrainingParams

__all__ = [
    "Actor", "WorldModelActor", "DenseModel", "ValueCritic", "ContinuousRSSM", "DiscreteRSSM", "RSSMBase",
    "InternalStateContinuous", "InternalStateContinuousSequence", "InternalStateDiscrete",
    "InternalStateDiscreteSequence", "AdamOptim", "FreezeParameters", "ValueGradTrainer",
    "ValueGradTrainerLosses", "ImaginedRollouts", "WorldModel", "WorldModelTrainingParams", "set_precision",
]
