"""Mirror of tf2.5/scripts/model/unets/__init__.py:1-5."""
from . import modelio, network_blocks, networks  # noqa: F401
