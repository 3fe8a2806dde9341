"""CPU oracles for the VectorQuantizer hot path -- test infrastructure, never the product path."""
