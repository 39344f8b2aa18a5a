"""Host-side mirror of the reference package ``tf2.5/scripts/model`` (model/__init__.py:1-2):
``model.unets.networks.M1`` and ``model.losses.Focal`` keep the reference names and arguments."""
from . import initializers, losses, optimizers, regularizers, unets  # noqa: F401
