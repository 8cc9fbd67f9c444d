class _Exc:
    class InvalidGitRepositoryError(Exception):
        pass


exc = _Exc


class Repo:
    def __init__(self, *a, **k):
        raise exc.InvalidGitRepositoryError("git stub")
