class PrettyTable:
    def __init__(self,*a,**k): pass
