/* stand-in for VTK's vtkNew.h: see vtk_standin.h (test infrastructure) */
#include "vtk_standin.h"
