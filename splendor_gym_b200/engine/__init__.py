from .rules import apply_action, initial_state, is_terminal, legal_moves, winner
from .state import Card, Noble, PlayerState, SplendorState

__all__ = ["SplendorState", "PlayerState", "Card", "Noble", "legal_moves", "apply_action", "is_terminal", "winner", "initial_state"]
