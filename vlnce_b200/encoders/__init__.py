"""Parameter containers of the three encoders (instruction, RGB, depth); their arithmetic is sequenced by
``vlnce_b200.engine`` / ``vlnce_b200.trunk`` and runs in the CUDA library."""
