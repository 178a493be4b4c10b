class Env(object):
    metadata = {}
    spec = None
    render_mode = None

    def reset(self, seed=None, options=None):
        # the real Env.reset only seeds self.np_random, which the reference never reads
        return None

    def step(self, action):
        raise NotImplementedError

    def close(self):
        pass
