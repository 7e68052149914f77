"""CPU oracle of the path-tracing hot path -- TEST INFRASTRUCTURE ONLY (see rtref.h)."""
