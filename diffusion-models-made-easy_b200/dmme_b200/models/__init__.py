from . import ddpm
from . import iddpm
