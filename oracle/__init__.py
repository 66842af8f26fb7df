"""CPU oracle for the pkg/despair SAD disparity path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package.  See oracle/sad_oracle.c for the reference file:line map.
"""
from .oracle import (  # noqa: F401
    build, lib, sum_abs_diff, region_literal, frame_box, frame_literal_mt,
    run_sad_chunks, output_camera_chunks, assemble_disparity_map,
)
