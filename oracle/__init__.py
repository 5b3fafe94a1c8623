"""CPU oracle of the reference hot path. TEST INFRASTRUCTURE ONLY (see module headers)."""
