"""No-op stand-in for yapf (only the reference's config pretty-printer imports it)."""


def FormatCode(text, **kw):
    return text, False
