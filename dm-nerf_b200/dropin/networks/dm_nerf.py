"""networks.dm_nerf (reference networks/dm_nerf.py): Embedder, get_embedder, DM_NeRF."""
from dmnerf_b200.embedder import Embedder, get_embedder   # noqa: F401
from dmnerf_b200.model import DM_NeRF                      # noqa: F401
