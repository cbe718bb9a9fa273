argscope = varreplace = optimizer = gradproc = None
