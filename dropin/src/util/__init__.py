"""Drop-in for the reference's `src/util` package: `import util; util.args.parse_args(...)`,
`util.gen_rays`, `util.pose_spherical`, ..."""
from pixel_nerf_multiscale_b200.util import *  # noqa: F401,F403
from pixel_nerf_multiscale_b200.util import args, conf  # noqa: F401
