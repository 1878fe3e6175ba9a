"""Import-time stand-in for gym==0.14.0 (absent from this image).

Test infrastructure only: lets the read-only reference under /root/reference be
imported in the build container to generate golden vectors.  Never imported by the
product package.
"""
