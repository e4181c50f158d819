"""networks.render (reference networks/render.py): dm_nerf, render_train."""
from dmnerf_b200.render import dm_nerf, render_train, render_rays, raw2outputs   # noqa: F401
