from typing import Tuple

NodeType = str
EdgeType = Tuple[str, str, str]
