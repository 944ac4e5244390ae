# Import-only stand-in: the reference imports pysam (phaser/phaser.py:15) but never calls it.
