class SyncServiceStub:
    def __init__(self, channel=None, **k): self.channel = channel
