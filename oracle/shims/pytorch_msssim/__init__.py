def ssim(*a, **k):
    raise NotImplementedError


class SSIM:
    def __init__(self, *a, **k):
        pass
