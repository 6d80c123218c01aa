"""Extra smoke checks appended as the hot path widens (called by __graft_entry__.smoke())."""


def run(pkg, ctx, orc):
    pass
