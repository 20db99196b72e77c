"""Model zoo: the architectures named by the reference's examples and BASELINE configs."""
