"""B200-native drop-in for the compress/decompress hot path of
TheJacksonLaboratory/cnn_autoencoder.  The names exported here are the ones the
reference's ``models`` package exports for that path
(``/root/reference/src/models/tasks/__init__.py``, ``criteria/__init__.py``).
"""
from ._autoencoders import (Analyzer, Synthesizer, ConvolutionalAutoencoder,
                            ConvolutionalAutoencoderBottleneck, autoencoder_from_state_dict,
                            setup_modules, load_state_dict, ModuleHandle)
from ._entropy import EntropyBottleneck
from ._taskutils import decorate_trainable_modules
from ._lossutils import GeneralLoss, RateLoss, DistMSELoss, setup_loss
from .train_step import (train_step, setup_optimizers, allreduce_gradients, GradBucket,
                         GraphedTrainStep)

__all__ = ['Analyzer', 'Synthesizer', 'ConvolutionalAutoencoder',
           'ConvolutionalAutoencoderBottleneck', 'autoencoder_from_state_dict', 'setup_modules',
           'load_state_dict', 'ModuleHandle', 'EntropyBottleneck', 'decorate_trainable_modules',
           'GeneralLoss', 'RateLoss', 'DistMSELoss', 'setup_loss', 'train_step',
           'setup_optimizers', 'allreduce_gradients', 'GradBucket', 'GraphedTrainStep']
