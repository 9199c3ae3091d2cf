"""CPU oracle of the mini-marl hot path -- test infrastructure, never on the product path."""
