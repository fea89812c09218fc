"""Card notation of the reference: two-character strings, rank in "23456789TJQKA", suit in "CDHS"
(reference tools/hand_evaluator.py:5-6).  Card id = 4*rank + suit = index in MonteCarlo.create_card_deck()
(reference tools/montecarlo_python.py:114-119)."""
import numpy as np

CARD_RANKS_ORIGINAL = "23456789TJQKA"
SUITS_ORIGINAL = "CDHS"
DECK = [r + s for r in CARD_RANKS_ORIGINAL for s in SUITS_ORIGINAL]
_CARD_ID = {c: i for i, c in enumerate(DECK)}
NO_CARD = 0xFF
HAND_TYPES = ["HighCard", "Pair", "TwoPair", "ThreeOfAKind", "Straight", "Flush", "FullHouse", "FoufOfAKind",
              "StraightFlush"]  # spelling as in the reference (hand_evaluator.py:92-115)


def card_id(card):
    """'AS' -> 51.  Raises ValueError for anything that is not a deck card, like list.index in the reference
    (montecarlo_python.py:127-128)."""
    try:
        return _CARD_ID[card]
    except (KeyError, TypeError):
        raise ValueError("%r is not in list" % (card,)) from None


def card_str(cid):
    return DECK[int(cid)]


def card_ids(cards):
    return [c if isinstance(c, (int, np.integer)) else card_id(c) for c in cards]


def encode_query(player_cards, table_cards):
    """(hole[2], board[5] padded with 0xFF) as uint8 arrays; validates like the reference would fail (ValueError)."""
    hole = card_ids(player_cards)
    board = card_ids(table_cards)
    if len(hole) != 2:
        raise ValueError("player_cards must hold exactly two cards, got %d" % len(hole))
    if len(board) > 5:
        raise ValueError("table_cards holds more than five cards")
    seen = set(hole) | set(board)
    if len(seen) != 2 + len(board):
        raise ValueError("duplicate cards in player_cards / table_cards")
    return (np.array(hole, dtype=np.uint8),
            np.array(board + [NO_CARD] * (5 - len(board)), dtype=np.uint8))
