/* empty stand-in for the system OpenGL header (glew.h includes it on non-Windows
 * builds); the oracle harness never calls GL.  TEST INFRASTRUCTURE. */
