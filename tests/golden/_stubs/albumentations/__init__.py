class DualTransform: 
    def __init__(self,*a,**k): pass
class RandomScale(DualTransform): pass
class ImageOnlyTransform(DualTransform): pass
def __getattr__(n): return DualTransform
