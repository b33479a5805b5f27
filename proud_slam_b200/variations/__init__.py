"""Host-side mirror of the reference's ``src/variations`` package for the render path:
``voxel_helpers`` (ray/octree intersection, sampling), ``render_helpers`` (feature lookup,
render_rays, the BA / tracking loops) and ``nrgbd`` (the decoder)."""
