"""Parameter trees of the reference model (U_Net, custom_layers); execution lives in b200/."""
