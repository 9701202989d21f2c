"""B200-native force path for AstroGenesis2.0 (octree build, group density, Barnes-Hut walk with
in-walk SPH) behind the reference's Tree call surface.  See DESIGN.md / INTEGRATION.md."""
from . import build, capi, ics, shard, tree  # noqa: F401
from .tree import Context, MultiContext, Simulation, Tree, run_step  # noqa: F401
