"""Score models.  Importing the package registers ``'ncsnpp'`` (``models/utils.register_model``)."""
from . import utils
from . import ncsnpp
