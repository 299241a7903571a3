Rectangle = Circle = None
