from typing import Callable

Schedule = Callable[[float], float]
