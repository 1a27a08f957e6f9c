"""transhuman_b200 -- B200-native (sm_100a) implementation of TransHuman's
per-ray volumetric query path behind the reference's ``Renderer`` plugin surface.

``transhuman_b200.ops``       torch-tensor front end of the C ABI (include/transhuman_b200.h)
``transhuman_b200.renderer``  drop-in for lib/networks/renderer/if_clight_renderer.py
``transhuman_b200.mesh_renderer`` drop-in for lib/networks/renderer/if_mesh_renderer.py
``transhuman_b200.synth``     deterministic synthetic frames (no dataset / checkpoint offline)
"""
__version__ = "0.1"
