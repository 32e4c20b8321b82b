"""CPU oracle for the PS hot path -- test infrastructure, never imported by the product."""
