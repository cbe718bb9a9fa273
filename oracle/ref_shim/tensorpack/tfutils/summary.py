def add_moving_summary(*a, **k):
    pass
