class Video:
    def __init__(self, frames, fps):
        self.frames, self.fps = frames, fps
