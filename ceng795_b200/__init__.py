"""ceng795_b200 — a B200-native rendering core behind the ceng795 ray tracer's
`Scene::render_image` seam.  Only what the hot path needs:
csrc/ (CUDA kernels + the C ABI + the host builder) and the host-side mirror of
the reference's Scene interface."""
from ._lib import B200RTError, RenderOpts, RenderStats, SEM_BY_NAME  # noqa: F401
from .scene import HostScene, Scene, parse_xml  # noqa: F401
