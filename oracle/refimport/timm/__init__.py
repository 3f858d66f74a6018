"""Minimal stand-in for timm==0.4.5 (requirements.txt:5 of the reference), which is not installed here.

Test infrastructure only: lets `oracle/refimport/loader.py` import the reference's vendored `models/`
package (which star-imports `timm.data` constants and a few `timm.models.*` names) so that golden
vectors can be generated from the UNMODIFIED reference code in this container.
"""
__version__ = "0.4.5"
