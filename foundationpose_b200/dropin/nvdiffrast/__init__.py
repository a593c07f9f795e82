"""Stand-in for the nvdiffrast package: the rasteriser lives inside libfpose.so (csrc/fp_crop.cu)."""
