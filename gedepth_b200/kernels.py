"""placeholder - replaced below"""
def has(name):
    return False
