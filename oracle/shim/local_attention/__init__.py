"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Stand-in for the PyPI package ``local-attention==1.11.2`` (pinned by the reference in
``pyproject.toml:10``), which is not vendored under /root/reference and cannot be installed here
(no network).  It restates the published algorithm of the three classes the reference uses
(``l3ac/local_trans.py:23``) so that the *unmodified* reference can be imported in the build
container to (a) validate ``oracle/l3ac_oracle.py`` and (b) generate ``tests/golden``.

PARITY UNPINNED: this restatement could not be checked against the real wheel.
"""
from .transformer import DynamicPositionBias, FeedForward, LocalMHA, LocalAttention  # noqa: F401
