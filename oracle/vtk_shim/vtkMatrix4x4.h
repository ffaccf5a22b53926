/* stand-in for VTK's vtkMatrix4x4.h: see vtk_standin.h (test infrastructure) */
#include "vtk_standin.h"
