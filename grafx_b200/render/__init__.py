from .captured import CapturedRender  # noqa: F401
from .graph import render_grafx  # noqa: F401
from .parallel import gather_batch, render_grafx_sharded, shard_batch, shard_bounds  # noqa: F401
from .plan import RenderData, mixing_console_plan, plan_from_dict  # noqa: F401
