/* stand-in for VTK's vtkInformationVector.h: see vtk_standin.h (test infrastructure) */
#include "vtk_standin.h"
