from .core import Model, Saver
from .generator import Generator, GSkip
from .discriminator import Discriminator
from .model import SEGAN, WSEGAN, weights_init, wsegan_weights_init

__all__ = ["Model", "Saver", "Generator", "GSkip", "Discriminator", "SEGAN", "WSEGAN", "weights_init",
           "wsegan_weights_init"]
