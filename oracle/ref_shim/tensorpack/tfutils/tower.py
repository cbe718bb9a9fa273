class _Ctx:
    is_training = False


def get_current_tower_context():
    return _Ctx()
