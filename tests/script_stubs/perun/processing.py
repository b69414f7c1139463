def processDataNode(root, cfg=None, force_process=False):
    return root
