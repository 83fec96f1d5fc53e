"""Drop-in alias: `from breakfast import breakfast, console, cache` resolves to the B200-native
implementation in breakfast_b200, so code and tests written against rki-mf1/breakfast run
unchanged (reference package layout: src/breakfast/{__init__,breakfast,cache,console}.py)."""
import sys as _sys

from breakfast_b200 import __version__, breakfast, cache, console  # noqa: F401

for _name, _mod in (("breakfast", breakfast), ("cache", cache), ("console", console)):
    _sys.modules[f"{__name__}.{_name}"] = _mod
