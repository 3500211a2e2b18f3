"""TEST INFRASTRUCTURE: checkers for the rasterizer path (see oracle/raster_oracle.c header).
Nothing under semantic-gaussians_b200/ imports this package."""
