class SeedStream:
    def __init__(self, seed=None, salt=None):
        self.seed = seed

    def __call__(self):
        return self.seed
